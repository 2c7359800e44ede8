timeout 300 python -m pytest tests/test_gpu_offpolicy.py -q -x 2>&1 | tail -5
timeout 120 python tools/profile_td3.py time 2>&1 | tail -7 | cut -c1-200
