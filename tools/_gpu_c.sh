M="lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum"
for extra in "" "--param b200_l2_keep_mb=40" "--param b200_l2_keep_mb=70" "--param b200_prefetch_chunks=0"; do
  echo "== ncu n=1250000 $extra"
  timeout 300 ncu --metrics $M --clock-control none -k regex:dual_solve -c 3 --csv python bench.py --n 1250000 --steps 2 --warmup 1 --no-cpu --no-e2e --no-parity --param dual_maxeval=40 $extra 2>/dev/null | grep -E "dual_solve" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | paste - - - - - - | head -3
done
python tools/trace_solve.py run 1250000 ccsaq
head -6 gpurun_out/trace_ccsaq_1250000.txt | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
