#!/bin/bash
# round 2, call v: fast block matcher (VABSDIFF4) parity + timing A/B (tile rows, generic vs fast)
cd ${GRAFT_REPO_ROOT:-/root/repo}
tag=${1:-r2v}
mkdir -p gpurun_out
echo "== stereo tests (fast path default)"; timeout 900 python -X faulthandler -m pytest tests/test_gpu_stereo.py -x -q > gpurun_out/${tag}_stereo_tests.log 2>&1; tail -15 gpurun_out/${tag}_stereo_tests.log
echo "== stereo tests (generic path forced)"; BPVO_B200_STEREO_GENERIC=1 timeout 900 python -m pytest tests/test_gpu_stereo.py -x -q 2>&1 | tail -3
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stereo.py -x -q -k "golden or random and 1 or random and 4" 2>&1 | tail -4
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stereo.py -x -q -k "golden" 2>&1 | tail -4
echo "== timing"
for ty in 8 16 32; do echo "tile rows $ty"; BPVO_B200_STEREO_TILE_ROWS=$ty python scripts/profile_stereo.py kitti; done
echo generic; BPVO_B200_STEREO_GENERIC=1 python scripts/profile_stereo.py kitti
for ty in 16 32; do echo "1080p tile rows $ty"; BPVO_B200_STEREO_TILE_ROWS=$ty python scripts/profile_stereo.py 1080p; done
echo "== ncu full"; REPS=2 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_bm_match -c 1 -o gpurun_out/${tag}_bm_match_fast python scripts/profile_stereo.py kitti > /dev/null 2>&1; ls -la gpurun_out/${tag}_bm_match_fast.ncu-rep
