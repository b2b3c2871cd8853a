# Round-end evidence on ONE GPU: plain runs first (exit 0), then the ncu passes of the same commands.
set -x
python bench.py > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_plain.json 2> gpurun_out/f_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_topm -s 3 -c 1 -o gpurun_out/f_scan_full -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_scan.log 2>&1
ncu -i gpurun_out/f_scan_full.ncu-rep --page raw --csv > gpurun_out/f_scan_full_raw.csv 2>/dev/null
python tools/batch_time.py 1250000 1024 1024 100 > gpurun_out/f_batch_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_batch_launches.csv python tools/batch_time.py 1250000 1024 1024 100 > gpurun_out/f_ncu_batch_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_gemm2 -s 9 -c 1 -o gpurun_out/f_batch_full -f python tools/batch_time.py 1250000 1024 1024 100 > gpurun_out/f_ncu_batch.log 2>&1
ncu -i gpurun_out/f_batch_full.ncu-rep --page raw --csv > gpurun_out/f_batch_full_raw.csv 2>/dev/null
tail -3 gpurun_out/f_batch_plain.log
cat gpurun_out/f_bench_n1.json | cut -c1-400
