python tools/overlap_probe.py 2>&1 | tail -1
for so in libeddsa_b200/variants/lib_*.so; do LIBEDDSA_B200_SO=$PWD/$so python tools/overlap_probe.py 2>&1 | tail -1; done
