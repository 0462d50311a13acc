python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hop_size" 2>&1 | tail -25
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "c1_click or c2_tracks or intermediates or escalation or accepted_config or hpss" 2>&1 | tail -3
