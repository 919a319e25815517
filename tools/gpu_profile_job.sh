set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_f64_tma -s 3 -c 1 -o gpurun_out/prof_bench_gemm_f64 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
python tools/quick_bench.py 16384,64,0.01,1e-6,spamm,f32 > gpurun_out/qb_f32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gemm_f32_tc -s 2 -c 1 -o gpurun_out/prof_gemm_f32_b64 -f python tools/quick_bench.py 16384,64,0.01,1e-6,spamm,f32 > gpurun_out/ncu_full_f32.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_1gpu.json | cut -c1-300; cat gpurun_out/bench_ref.json | cut -c1-600; cat gpurun_out/qb_f32.log; ls -la gpurun_out | tail -12
