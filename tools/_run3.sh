mkdir -p gpurun_out
for m in 0 4; do echo "LMB_RAY_KEY=$m"; LMB_RAY_KEY=$m python tools/torus_sweep.py --max-log2 24 --out gpurun_out/torus_key$m.json 2>&1 | grep "sorted" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['rays'], round(d['mrays_per_s']), round(d['ms_per_launch'],3))"; done
nsys --version 2>/dev/null | head -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/torus_launches.csv python tools/torus_sweep.py --max-log2 24 --out gpurun_out/torus_ncu.json > /dev/null 2>&1
grep -E "k_ray_keys|k_ray_scatter|k_scan_exclusive|k_trace_array" gpurun_out/torus_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -30
