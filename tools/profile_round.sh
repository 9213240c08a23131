# The round's profile pass (under gpurun, one GPU): per-kernel metrics of one render per configuration, --set full
# captures of the mesh pipeline's kernels and of the endgame, and the launch list of the bench.
bash tools/ncu_all.sh
python tools/one_render.py C3 10 1 > /dev/null && SHIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on \
  -k regex:"wf_bvh1_list|wf_bvh1_walk|wf_bvh1_finish|wf_shade" -s 4 -c 4 -o gpurun_out/r2_mesh -f python tools/one_render.py C3 10 1 > gpurun_out/r2_mesh.log 2>&1
python tools/one_render.py C5 2 1 predictor=False > /dev/null && SHIM_NO_GRAPH=1 ncu --set full --clock-control none --import-source on \
  -k regex:"wf_bvh1_walk" -s 1 -c 1 -o gpurun_out/r2_mesh_c5 -f python tools/one_render.py C5 2 1 predictor=False > gpurun_out/r2_mesh_c5.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-parity --e2e-steps 0 --profile-steps 0 > gpurun_out/r2_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --no-parity --e2e-steps 0 --profile-steps 0 > gpurun_out/r2_ncu.log 2>&1
