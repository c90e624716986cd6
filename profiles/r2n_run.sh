set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" 
tail -5 gpurun_out/r2n_tests.log
VARIANTS="B C" bash profiles/ab.sh 2>&1 | tee gpurun_out/r2n_ab.log
cp profiles/ab/C.so jpeg-encoder-opencl_b200/libjpegb200.so
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 240 --csv --log-file gpurun_out/r2n_launches.csv \
    -k 'regex:k_transform|k_fixup|k_encode|k_scan|k_intervals|k_zero|k_pack|k_ff_count|k_int_out|k_finalize|k_stuff' $B > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2n_launches.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr=rows[hi]; kn=hdr.index('Kernel Name'); mv=hdr.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>mv: d[r[kn].split('(')[0]].append(float(r[mv].replace(',','')))
for k,v in d.items(): print(k, len(v), 'avg us %.1f'%(sum(v)/len(v)/1000))
PY
ncu --set full --import-source on --clock-control none -k 'regex:k_encode|^k_pack$|k_ff_count|^k_stuff$' \
    --launch-skip 12 --launch-count 4 -f -o gpurun_out/r2n_prof_ent $B > /dev/null 2>&1
ls -la gpurun_out/r2n_*
