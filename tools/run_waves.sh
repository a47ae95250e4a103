for w in 8 14 16; do EDDSA_B200_VERIFY_WAVES=$w python tools/tune_bench.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('waves', $w, 'verify', d['verify'], d['ok'])"; done
bash tools/run_variants.sh
