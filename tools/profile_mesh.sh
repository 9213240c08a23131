set -x
cap() { cfg=$1; label=$2; spp=$3; shift 3; python tools/one_render.py $cfg $spp 2 "$@" > gpurun_out/r2_ncu_$label.out 2>gpurun_out/r2_ncu_$label.err && SHIM_NO_GRAPH=1 ncu --metrics $(python tools/ncu_summary.py --metrics) --clock-control none --csv --log-file gpurun_out/r2_ncu_$label.csv python tools/one_render.py $cfg $spp 1 "$@" > /dev/null 2>&1; tail -1 gpurun_out/r2_ncu_$label.out; }
cap C3 C3 4
cap C5 C5 1 predictor=False
python tools/one_render.py C3 10 1 > /dev/null && SHIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on \
  -k regex:"wf_bvh1_list|wf_bvh1_walk|wf_bvh1_finish|wf_shade" -s 4 -c 4 -o gpurun_out/r2_mesh -f python tools/one_render.py C3 10 1 > gpurun_out/r2_mesh.log 2>&1
python tools/one_render.py C5 2 1 predictor=False > /dev/null && SHIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on \
  -k regex:"wf_bvh1_walk" -s 1 -c 1 -o gpurun_out/r2_mesh_c5 -f python tools/one_render.py C5 2 1 predictor=False > gpurun_out/r2_mesh_c5.log 2>&1
