# C3 (VARCHAR-heavy) through the pack kernel with different stage sizings; one process per variant (the knobs are read once)
for v in "" "DMB_STR_PACK_HPR_LIMIT=40" "DMB_STR_PACK_HPR_LIMIT=40 DMB_STR_PACK_SLACK=1.05" "DMB_STR_PACK_HPR_LIMIT=40 DMB_STR_PACK_SLACK=1.0" "DMB_STR_PACK_HPR_LIMIT=40 DMB_STR_PACK_SLACK=1.3" "DMB_STR_PACK_HPR_LIMIT=40 DMB_STR_PACK_SLACK=1.05 DMB_STR_PACK_CTAS=2"; do
  echo "== $v"; env $v python profiles/bench_configs.py --configs c3 --scale 0.5 2>&1 | tail -1 | cut -c1-330
done
