set -x
cap() { cfg=$1; label=$2; spp=$3; shift 3; python tools/one_render.py $cfg $spp 2 "$@" > gpurun_out/r2_ncu_$label.out 2>gpurun_out/r2_ncu_$label.err && SHIM_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2_ncu_$label.csv python tools/one_render.py $cfg $spp 1 "$@" > /dev/null 2>&1; tail -1 gpurun_out/r2_ncu_$label.out; }
cap C1 C1 10
cap C2 C2 16
cap C3 C3 4
cap C4 C4 8 predictors=False
cap C4 C4-hrpp 8 predictors=True
cap C5 C5 1 predictor=False
cap C5 C5-hrpp 1 predictor=True
