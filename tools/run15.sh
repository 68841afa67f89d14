for bn in 128 256; do echo "DW_BN=$bn"; MLB_TF32_DW_BN=$bn timeout 300 python tools/tf32_bench.py 2>&1 | grep -E "dW|cfg2" | cut -c1-170; done
