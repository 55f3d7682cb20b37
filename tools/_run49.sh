python tools/grid3d_demo.py 63 2 2>/dev/null | tail -1 > gpurun_out/grid3d_demo_r2_63.json; cat gpurun_out/grid3d_demo_r2_63.json | cut -c1-900
python tools/grid3d_demo.py 127 2 2>/dev/null | tail -1 > gpurun_out/grid3d_demo_r2_127.json; cat gpurun_out/grid3d_demo_r2_127.json | cut -c1-900
