cd "${GRAFT_REPO_ROOT:-/root/repo}"; mkdir -p gpurun_out
for w in mri_r10 mri_r18 pet_mri_fusion_faithful all_modalities mri_r50_160; do
  timeout 500 python bench.py --workload $w --steps 4 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"; python - gpurun_out/bench_$w.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value'],1),'vol/s',round(d['ms_per_step'],3),'ms','e2e',round(d['e2e']['value'],1),'cpu',round(d['cpu_baseline']['value'],2),d['clocks'],'frac',round(d['roofline']['frac'] or 0,3),round(d['roofline']['whole_step_tensor_frac'],3))
except Exception as e: print('no line',e)
PY
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"; cut -c1-300 gpurun_out/bench_reference.json
