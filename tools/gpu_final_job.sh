set -x
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_1gpu.err | grep "^{" > gpurun_out/bench_1gpu.json && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | grep "^{" > gpurun_out/bench_ref.json
python tools/quick_bench.py 32768,32,0.02,1e-6 32768,128,0.01,1e-6 32768,256,0.01,1e-6 16384,64,0.01,1e-6,spamm,f32 > gpurun_out/qb_sweep.log 2>&1
python tools/hbm_stages.py > gpurun_out/hbm_stages.jsonl 2>/dev/null
cut -c1-250 gpurun_out/bench_1gpu.json; cut -c1-200 gpurun_out/bench_ref.json; cat gpurun_out/qb_sweep.log | cut -c1-420
