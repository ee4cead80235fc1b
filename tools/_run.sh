timeout 1000 python -m pytest tests -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py > gpurun_out/r5b_bench_n1.json 2> gpurun_out/r5b_bench_n1.err
echo rc=$?; tail -3 gpurun_out/r5b_bench_n1.err; wc -c gpurun_out/r5b_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
