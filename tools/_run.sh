timeout 200 python -m pytest tests/test_block_kernels.py tests/test_parity_quant.py -q -m gpu -k "act_quant or repeat_kv or backward" --tb=short > gpurun_out/r5c_new_tests.log 2>&1
tail -40 gpurun_out/r5c_new_tests.log
timeout 600 python -m pytest tests -q -m gpu --tb=line > gpurun_out/r5c_all_tests.log 2>&1
tail -25 gpurun_out/r5c_all_tests.log
timeout 300 python bench.py > gpurun_out/r5c_bench_n1.json 2> gpurun_out/r5c_bench_n1.err
echo rc=$?; tail -3 gpurun_out/r5c_bench_n1.err; wc -c gpurun_out/r5c_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
