python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
python tools/perf_probe.py 708 1920 1080 instanced 2>&1 | tail -42 > gpurun_out/probe_c3.txt; grep -E "render_Mrays|render_ms|primary_Mrays|incoherent_Mrays|upload|rays_per|ms_extend|ms_connect|ms_shade" gpurun_out/probe_c3.txt
python tools/perf_probe.py 708 1920 1080 motion 2>&1 | tail -42 > gpurun_out/probe_c4.txt; grep -E "render_Mrays|render_ms|primary_Mrays|incoherent_Mrays|upload|rays_per|ms_extend|ms_connect|ms_shade" gpurun_out/probe_c4.txt
