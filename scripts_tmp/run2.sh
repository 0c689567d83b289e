timeout 60 python scripts_tmp/dbg_mlp.py 2>&1 | tail -3
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python profiles/microbench_gemm.py 10 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; tail -1 gpurun_out/bench4.json | cut -c1-200
