# Round-2 evidence on ONE GPU: plain runs first (exit 0), then the ncu passes of the same commands.
# Everything lands in gpurun_out/p2_*; copy the summaries into profiles/ (see profiles/README.md).
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/p2_bench_n1.json 2> gpurun_out/p2_bench_n1.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/p2_bench_ref.json 2> gpurun_out/p2_bench_ref.err
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none"
$CMD > gpurun_out/p2_plain.json 2> gpurun_out/p2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/p2_launches.csv $CMD > gpurun_out/p2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_topm -s 3 -c 1 -o gpurun_out/p2_scan_full -f $CMD > gpurun_out/p2_ncu_scan.log 2>&1
ncu -i gpurun_out/p2_scan_full.ncu-rep --page raw --csv > gpurun_out/p2_scan_full_raw.csv 2>/dev/null
python tools/make_traffic.py gpurun_out/p2_scan_full_raw.csv 10000000 768 gpurun_out/p2_traffic.json
# batched contraction, three operand precisions (config-4 shape per GPU)
python tools/batch_time.py 1250000 1024 1024 100 > gpurun_out/p2_batch_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/p2_batch_launches.csv python tools/batch_time.py 1250000 1024 1024 100 > gpurun_out/p2_ncu_batch_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:batch_gemm2 -s 9 -c 1 -o gpurun_out/p2_batch_full -f python tools/batch_time.py 1250000 1024 1024 100 > gpurun_out/p2_ncu_batch.log 2>&1
ncu -i gpurun_out/p2_batch_full.ncu-rep --page raw --csv > gpurun_out/p2_batch_full_raw.csv 2>/dev/null
# latency path (config 1): at the C ABI (no interpreter), then through ctypes, then the in-kernel timeline
gcc -O2 -Iinclude tools/lat_bench.c -o /tmp/lat_bench -Lrust-local-rag_b200 -l:librlr_b200.so -Wl,-rpath,$PWD/rust-local-rag_b200 -lm && \
  { /tmp/lat_bench 10000 5 0.3; /tmp/lat_bench 10000 5 0.3 0x20; /tmp/lat_bench 10000 10 0.3; /tmp/lat_bench 10000 100 0.3; /tmp/lat_bench 100000 5 0.3; } > gpurun_out/p2_lat_cabi.log 2>&1
python tools/lat_trace.py 10000 5 0.3 > gpurun_out/p2_lat_plain.log 2>&1
RLR_DEBUG_LAT_TRACE=1 python tools/lat_trace.py 10000 5 0.3 2>&1 | grep "lat trace" | sed -n "5,12p" > gpurun_out/p2_lat_trace.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 40 --csv --log-file gpurun_out/p2_lat_launches.csv python tools/lat_trace.py 10000 5 0.3 > gpurun_out/p2_ncu_lat.log 2>&1
rm -f gpurun_out/p2_scan_full.ncu-rep gpurun_out/p2_batch_full.ncu-rep
tail -3 gpurun_out/p2_batch_plain.log; cat gpurun_out/p2_lat_cabi.log gpurun_out/p2_lat_plain.log; cut -c1-300 gpurun_out/p2_bench_n1.json
