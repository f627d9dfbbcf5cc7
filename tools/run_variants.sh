for v in b4sb b2sb b4sbf4; do
  if [ -z "$v" ]; then unset SDB_LIBRARY; else export SDB_LIBRARY=$PWD/scikit-downscale_b200/csrc/variants/libsdb_$v.so; fi
  echo -n "variant=$v "; python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
