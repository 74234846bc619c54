set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r2w_pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -12 gpurun_out/r2w_pytest_gpu.log
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
echo "bench rc $?"; tail -c 1500 gpurun_out/r2w_bench.json; tail -4 gpurun_out/r2w_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err
echo "ref rc $?"; tail -c 300 gpurun_out/r2w_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2w_ncu_block16.csv python tools/ncu_block.py 16 > gpurun_out/r2w_ncu_block16.out 2>&1
echo "ncu block rc $?"; python tools/ncu_summary.py gpurun_out/r2w_ncu_block16.csv gpurun_out/r2w_ncu_block16 | head -12
timeout 200 python tools/cm_prof2.py 1 > gpurun_out/r2w_cm_prof.log 2>&1; cat gpurun_out/r2w_cm_prof.log
