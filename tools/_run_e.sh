mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_swinir_gpu.py tests/test_wsti_gpu.py tests/test_parity_hardening_gpu.py -m gpu -q 2>&1 | tail -5) > gpurun_out/r02_ai_pytest.log 2>&1
cat gpurun_out/r02_ai_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_ai_bench_c3.log 2>gpurun_out/r02_ai_bench_c3.err; tail -c 200 gpurun_out/r02_ai_bench_c3.log; cp gpurun_out/kernel_table_n1.json gpurun_out/r02_ai_kernel_table_c3.json
