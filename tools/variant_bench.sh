# usage: variant_bench.sh "<-D flags>" ...   -- builds a variant library on the box and runs bench.py on it
for defs in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -ftz=true -std=c++17 -shared -Xcompiler -fPIC $defs -o /tmp/libv.so sq_recovery_b200/csrc/sqloss.cu || exit 1
  SQ_LIBSQLOSS=/tmp/libv.so python bench.py 2>/dev/null > /tmp/b.json
  python -c "
import json; d=json.load(open('/tmp/b.json')); print('$defs |', round(d['value']), round(d['ms_per_step']*1e3,1), round(d['one_batch_at_a_time']['ms_per_step']*1e3,1), round(d['roofline']['kernel_ms']*1e3,1), 'dense', round(d['dense']['value']), round(d['dense']['ms_per_step']*1e3,1), round(d['dense']['roofline']['kernel_ms']*1e3,1), round(d['dense']['roofline']['frac'],3))"
done
