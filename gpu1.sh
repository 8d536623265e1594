set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; lscpu | grep "Model name"
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -5
