for v in imad imadgr; do PMCTF_LIB=$PWD/learned-pmctf_b200/lib/libpmctf_b200_$v.so timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gop.py -x -q -m gpu -k "predict_update or mctf or lift or config2 or concurrent" 2>&1 | tail -1; done
bash tools/ab_variants.sh r2w11 imad imadgr
timeout 100 python -m pytest tests/test_gpu_llar.py -x -q -m gpu 2>&1 | tail -1
