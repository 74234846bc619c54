set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > gpurun_out/r3d_pytest_gpu.log 2>&1
echo "pytest rc $?"; tail -12 gpurun_out/r3d_pytest_gpu.log
timeout 300 python tools/inflight_curve.py --mib 4 --ks 1,6,37,74,128,148,296 --out gpurun_out/r3d_inflight_zipf4m.json > gpurun_out/r3d_inflight_zipf4m.log 2>&1; tail -12 gpurun_out/r3d_inflight_zipf4m.log
( time timeout 1500 python bench.py --steps 3 --warmup 3 ) > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
echo "bench rc $?"; tail -c 400 gpurun_out/r3d_bench.json; tail -4 gpurun_out/r3d_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r3d_bench_ref.json 2> gpurun_out/r3d_bench_ref.err
echo "ref rc $?"; tail -c 300 gpurun_out/r3d_bench_ref.json
