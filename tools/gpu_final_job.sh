set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_1gpu.err > gpurun_out/bench_1gpu.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err > gpurun_out/bench_ref.json
timeout 600 python tools/run_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python tools/hbm_stages.py > gpurun_out/hbm_stages.jsonl 2>/dev/null
cut -c1-250 gpurun_out/bench_1gpu.json; cut -c1-200 gpurun_out/bench_ref.json; tail -3 gpurun_out/configs.err
