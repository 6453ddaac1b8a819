# pass-1 variants of the 16-mer path (built with findkmer_b200.build.build_variant)
for v in "" _w640 _w_skel; do
  echo "variant${v}"
  FKB_LIB=findkmer_b200/libfindkmer_b200${v}.so timeout 200 python profiles/tools/phase_probe.py 4 3100000000
done
