python tools/dec_attn_trace.py points_api 2>&1 | tail -10
python -m pytest tests/test_engine_gpu.py tests/test_image_predictor.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -5
