#!/bin/bash
# compute-sanitizer passes over the whole path on a small scene (smoke-sized): memcheck, racecheck (shared memory), synccheck
cd ${GRAFT_REPO_ROOT:-/root/repo}
mkdir -p gpurun_out
cat > /tmp/san_run.py <<'P'
import numpy as np, sys
sys.path.insert(0, '.')
from bpvo_b200 import AlgorithmParameters, DescriptorType, LossFunctionType, VerbosityType, VisualOdometry, synth
from bpvo_b200.types import InterpolationType
for desc, loss, interp, sct in [(DescriptorType.kBitPlanes, LossFunctionType.kTukey, InterpolationType.kLinear, -1.0),
                                (DescriptorType.kIntensity, LossFunctionType.kHuber, InterpolationType.kCubic, -1.0),
                                (DescriptorType.kBitPlanes, LossFunctionType.kL2, InterpolationType.kLinear, 0.75)]:
    sc = synth.scene_small(96, 128)
    p = AlgorithmParameters(descriptor=desc, numPyramidLevels=2, lossFunction=loss, verbosity=VerbosityType.kSilent, interp=interp,
                            sigmaPriorToCensusTransform=sct, maxIterations=6)
    vo = VisualOdometry(sc.K, sc.baseline, (sc.rows, sc.cols), p)
    for k in range(3):
        r = vo.addFrame(*sc.render(k))
    print("ok", int(desc), int(loss), int(interp), r.numFunEvals, flush=True)
    # fine seam too
    ctx = vo.ctx
    vo.close()
print("done")
P
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_run.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|done|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
