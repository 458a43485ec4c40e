timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/r2p_all.log 2>&1; echo "gpu tests exit $?"; tail -3 gpurun_out/r2p_all.log
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | cut -c1-200
