TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29804 tools/config_runs.py --config 4 --steps 20 2>&1 | grep '^{' > gpurun_out/i_config4_n8.json
timeout 300 $TR --nproc-per-node 8 --master-port 29805 tools/config_runs.py --config 4 --steps 20 --exact 2>&1 | grep '^{' > gpurun_out/i_config4_n8_exact.json
cat gpurun_out/i_config4_n8.json gpurun_out/i_config4_n8_exact.json | cut -c1-900
