set -x
cp profiles/ab/C.so jpeg-encoder-opencl_b200/libjpegb200.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?" 
tail -15 gpurun_out/r2ac_tests.log
VARIANTS="B C" bash profiles/ab.sh 2>&1 | tee gpurun_out/r2ac_ab.log
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others --no-parity"
for v in C; do
cp profiles/ab/$v.so jpeg-encoder-opencl_b200/libjpegb200.so
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 240 --csv --log-file gpurun_out/r2ac_launches_$v.csv \
    -k 'regex:k_encode|k_zero|k_pack|k_ff_count|k_stuff|k_fixup' $B > /dev/null 2>&1
python - $v <<'PY'
import csv,collections,sys
rows=list(csv.reader(open('gpurun_out/r2ac_launches_%s.csv'%sys.argv[1])))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value'); mn=hdr.index('Metric Name'); mu=hdr.index('Metric Unit')
d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>mv: d[(r[kn].split('(')[0], r[mn], r[mu])].append(float(r[mv].replace(',','')))
for k,v in d.items(): print(sys.argv[1], k, len(v), 'avg %.1f'%(sum(v)/len(v)))
PY
done
