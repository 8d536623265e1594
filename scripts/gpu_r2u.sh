#!/bin/bash
# round 2, call u: stereo tests with full logs, ncu of the block matcher
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2u}
mkdir -p gpurun_out
echo "== stereo tests"; timeout 900 python -X faulthandler -m pytest tests/test_gpu_stereo.py -x -q > gpurun_out/${tag}_stereo_tests.log 2>&1; head -c 3000 gpurun_out/${tag}_stereo_tests.log | head -40; tail -5 gpurun_out/${tag}_stereo_tests.log
echo "== sanitizer (random configurations)"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stereo.py -x -q -k "random and 0 or random and 3" 2>&1 | tail -6
echo "== plain timing"; python scripts/profile_stereo.py kitti; python scripts/profile_stereo.py 1080p
echo "== ncu launches"; REPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_stereo_launches.csv python scripts/profile_stereo.py kitti > /dev/null 2>&1; tail -4 gpurun_out/${tag}_stereo_launches.csv
echo "== ncu full"; REPS=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_bm_match -c 1 -o gpurun_out/${tag}_bm_match python scripts/profile_stereo.py kitti > /dev/null 2>&1; ls -la gpurun_out/${tag}_bm_match.ncu-rep
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
