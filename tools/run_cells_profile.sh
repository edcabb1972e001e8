set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cell_g1_fft_stage -s 45 -c 2 -o gpurun_out/d3_cell_fft python tools/cells_bench.py 864 > gpurun_out/d3_ncu_fft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msm_gather_ba -s 2 -c 1 -o gpurun_out/d3_cell_ba python tools/cells_bench.py 864 > gpurun_out/d3_ncu_ba.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/d3_cells_launches_864.csv python tools/cells_bench.py 864 > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_cells.py -x -q -k "kats or edge or batch_and_device or verify_cell or recover or large_batch" > gpurun_out/d3_sanitizer_cells.log 2>&1
tail -15 gpurun_out/d3_sanitizer_cells.log
ls -la gpurun_out/d3_*
