set -x
nvidia-smi --query-gpu=name --format=csv,noheader
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -s 2>&1 | tail -40
