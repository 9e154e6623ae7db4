# Historical (round 1): the JG_MS_VARIANT switch (ticket vs static tile order, look-back depth 1 vs 8) existed only
# while the single-pass look-back mask_select was being studied; results are in r1_lookback_study.md section 2.
for v in 0 1 2 3; do
  JG_MS_VARIANT=$v python bench.py --no-cpu --steps 2 --comb-events 1000000 > gpurun_out/ms_v$v.json 2>gpurun_out/ms_v$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ms_v$v.json'))
print($v, d['ops']['mask_select_f32']['ms'], d['ops']['segargmax_dense_f32']['ms'], d['ops']['counts2offsets']['ms'])
PY
done
