for ty in 64 500; do timeout 120 python tools/train_forward_probe.py --batch 4 --ty $ty --reps 3 2>/dev/null | cut -c1-200; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file /tmp/tf.csv python tools/train_forward_probe.py --batch 1 --reps 1 > /dev/null 2>&1
python - <<PY
import csv,re,collections
rows=[r for r in csv.reader(open("/tmp/tf.csv")) if len(r)>5]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=re.sub(r"\(.*","",r[ki].replace("(int)","").replace("(bool)",""))[:60]; agg[n][0]+=1; agg[n][1]+=float(r[vi].replace(",",""))/1e6
tot=sum(v for _,v in agg.values()); print("kernel ms total (2 forwards of 1 utterance + load)", round(tot,2))
for n,(c,v) in sorted(agg.items(), key=lambda z:-z[1][1])[:8]: print(round(v,3), c, n)
PY
