timeout 600 python tools/pucci_scaling.py 6 255 1023 2047 4095
timeout 300 python tools/pucci_scaling.py 2 255 1023
