# timing experiment: grid.x (parts per image) of the first-generation gm_likelihood_kernel (NIC_LIK_FLAT=0) at batch 16 / 256;
# the default since late round 2 is the flat kernel - tools/lik_bench.py compares the two
for p in 27 37 46 55 64; do
  NIC_LIK_FLAT=0 NIC_LIK_PARTS=$p python bench.py --no-train-step --no-scalable --no-other-arms --no-cpu-baseline --steps 3 2>/dev/null > /tmp/l.json
  python -c "
import json; d=json.load(open('/tmp/l.json')); r=d['roofline_likelihood']; print($p, round(r['batch16']['ms']*1e3,1), round(r['batch16']['frac_of_hbm_peak'],3), round(r['batch256']['frac_of_hbm_peak'],3))"
done
