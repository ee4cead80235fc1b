for l in "" _tca_650c23b _tca_3e478c0; do echo "== lib$l"; Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100$l.so timeout 200 python tools/exp_tc_shapes.py 2>&1 | tail -5; done
