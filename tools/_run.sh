timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/r5c_all_tests.log 2>&1
grep -v "^[.sF]*  *\[" gpurun_out/r5c_all_tests.log | tail -60
timeout 300 python bench.py > gpurun_out/r5c_bench_n1.json 2> gpurun_out/r5c_bench_n1.err
echo rc=$?; tail -3 gpurun_out/r5c_bench_n1.err; wc -c gpurun_out/r5c_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
