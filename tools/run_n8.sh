# 8-GPU evidence run (one box): fused-exchange parity, the bench line, BASELINE configs 4 and 5.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
RLR_DIST_MODE=fused timeout 200 $TR --nproc-per-node 8 --master-port 29801 tools/dist_parity.py 150001 768 2>&1 | grep DIST_PARITY > gpurun_out/g_parity8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 200 --warmup 10 > gpurun_out/g_bench_n8.json 2> gpurun_out/g_bench_n8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29804 tools/config_runs.py --config 4 --steps 10 2>&1 | grep '^{' > gpurun_out/g_config4_n8.json
timeout 300 $TR --nproc-per-node 8 --master-port 29805 tools/config_runs.py --config 4 --steps 10 --exact 2>&1 | grep '^{' > gpurun_out/g_config4_n8_exact.json
timeout 400 $TR --nproc-per-node 8 --master-port 29806 tools/config_runs.py --config 5 --steps 50 2>&1 | grep '^{' > gpurun_out/g_config5_n8.json
cat gpurun_out/g_parity8.log gpurun_out/g_bench_n8.json gpurun_out/g_config4_n8.json gpurun_out/g_config4_n8_exact.json gpurun_out/g_config5_n8.json | cut -c1-1600
tail -3 gpurun_out/g_bench_n8.err
