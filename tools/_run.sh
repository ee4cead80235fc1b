timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r4i_bench_n1.json 2> gpurun_out/r4i_bench_n1.err
echo rc=$?; tail -2 gpurun_out/r4i_bench_n1.err; wc -c gpurun_out/r4i_bench_n1.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
