mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_b.log 2>&1
python - <<'PY'
import csv,collections
t=collections.defaultdict(lambda:[0,0])
for r in csv.reader(open('gpurun_out/launches_tc.csv')):
    if len(r)>10 and r[0].isdigit():
        n=r[4].split('(')[0][:70]; t[n][0]+=float(r[-1]); t[n][1]+=1
tot=sum(v[0] for v in t.values())
for n,v in sorted(t.items(), key=lambda kv:-kv[1][0])[:24]: print(f"{100*v[0]/tot:5.1f}%  {v[0]/v[1]/1e3:8.1f} us x{v[1]:4d}  {n}")
PY
