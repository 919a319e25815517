set -x
python tools/hbm_stages.py > gpurun_out/hbm_stages.jsonl 2> gpurun_out/hbm_stages.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_leaf_norms|k_add_tiles|k_transpose_tiles|k_join|k_rs_scatter|k_task_finish" -c 60 --csv --log-file gpurun_out/hbm_kernels.csv python tools/hbm_stages.py > gpurun_out/ncu_hbm.log 2>&1
timeout 1200 python tools/run_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
cat gpurun_out/hbm_stages.jsonl; tail -2 gpurun_out/hbm_stages.err; tail -3 gpurun_out/configs.err
