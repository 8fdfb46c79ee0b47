"""Encoding several frames per pass of the image encoder (predictor._get_image_feature / engine.encode_images).

The backbone features of a frame depend on that frame only (sam2_video_predictor.py:1174-1212), so encoding the next
few frames of the processing order together must not change anything the predictor produces.  CPU: the predictor's
bookkeeping on the oracle engine (forward, reverse with a window, feature cache on / off, frames released).  GPU:
``encode_images`` is bit-identical per frame to ``encode_image`` and a tracked sequence is bit-identical with and
without it.
"""
import numpy as np
import pytest
import torch

from detsam2_b200.config import get_config
from detsam2_b200.predictor import SAM2VideoPredictor
from detsam2_b200.synthetic import BilliardVideo
from detsam2_b200.weights import synthetic_state_dict


def _track(pred, vid, nframes, reverse=False, chunks=1):
    out = []
    with torch.inference_mode():
        per = nframes // chunks
        st = pred.init_state([vid.frame(t) for t in range(per)])
        for oid, b in vid.boxes(0).items():
            pred.add_new_points_or_box(st, 0, oid, box=np.asarray(b, dtype=np.float32))
        if not reverse:
            out += [(f, list(i), m.float().cpu().clone()) for f, i, m in pred.propagate_in_video(st)]
        for c in range(1, chunks):
            st = pred.update_state([vid.frame(t) for t in range(c * per, (c + 1) * per)], st)
            last = (c + 1) * per - 1
            if reverse:
                for oid, b in vid.boxes(c * per).items():     # Det-SAM2 prompts every chunk (detect_interval = K)
                    pred.add_new_points_or_box(st, c * per, oid, box=np.asarray(b, dtype=np.float32))
                out += [(f, list(i), m.float().cpu().clone())
                        for f, i, m in pred.propagate_in_video(st, start_frame_idx=last, max_frame_num_to_track=2 * per,
                                                               reverse=True)]
                pred.release_old_frames(st, last, per, 0, release_images=True)
            else:
                out += [(f, list(i), m.float().cpu().clone())
                        for f, i, m in pred.propagate_in_video(st, start_frame_idx=c * per)]
    return out, st


def _same(a, b, exact):
    assert len(a) == len(b)
    for (fa, ia, ma), (fb, ib, mb) in zip(a, b):
        assert fa == fb and ia == ib
        if exact:
            assert torch.equal(ma, mb), fa
        else:
            assert torch.allclose(ma, mb, atol=1e-5), fa


@pytest.mark.parametrize("reverse,chunks,cache", [(False, 1, 1), (False, 2, 1), (True, 3, 1), (True, 3, 16)])
def test_encode_ahead_bookkeeping_on_oracle_engine(reverse, chunks, cache):
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=256)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=2, height=96, width=128, num_frames=9, seed=3)

    def run(E):
        eng = O.OracleEngine(cfg, sd, fill_holes=False)
        pred = SAM2VideoPredictor(eng, fill_hole_area=0, feature_cache_frames=cache, encoder_batch_frames=E)
        res, st = _track(pred, vid, 9 if chunks > 1 else 6, reverse=reverse, chunks=chunks)
        assert pred._upcoming is None and pred._prefetched == {}        # nothing outlives the propagate call
        return res, getattr(eng, "encode_images_calls", 0), st

    one, calls1, _ = run(1)
    four, calls4, st4 = run(4)
    assert calls1 == 0 and calls4 > 0
    _same(one, four, exact=True)           # the oracle encodes frame by frame either way: identical arithmetic
    assert len(st4["images"]) == len(st4["images_idx"])


def test_encode_ahead_abandoned_generator_leaves_no_state():
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=256)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=1, height=96, width=128, num_frames=6, seed=4)
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, sd, fill_holes=False), fill_hole_area=0, encoder_batch_frames=4)
    with torch.inference_mode():
        st = pred.init_state([vid.frame(t) for t in range(6)])
        pred.add_new_points_or_box(st, 0, 0, box=np.asarray(vid.boxes(0)[0], dtype=np.float32))
        gen = pred.propagate_in_video(st)
        next(gen)
        next(gen)
        assert pred._prefetched            # frames 2.. were encoded ahead
        gen.close()
        assert pred._upcoming is None and pred._prefetched == {}
        # a later prompt on another frame encodes on its own
        pred.add_new_points_or_box(st, 5, 0, box=np.asarray(vid.boxes(5)[0], dtype=np.float32))


@pytest.mark.parametrize("reverse,chunks,cache", [(False, 1, 1), (False, 2, 1), (True, 3, 1), (True, 3, 16)])
def test_overlapped_encoder_bookkeeping_on_oracle_engine(reverse, chunks, cache):
    """The pass over the NEXT frames is launched (engine.encode_images_async) when the current pass starts to be
    consumed and joined when its first frame is reached: same results, every frame encoded exactly as often as without
    the overlap, at most one pass in flight, nothing left behind."""
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=256)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=2, height=96, width=128, num_frames=12, seed=3)

    class AsyncOracle(O.OracleEngine):
        launched = joined = frames_encoded = 0
        pred = None

        def encode_image(self, image):
            self.frames_encoded += 1
            return super().encode_image(image)

        def encode_images_async(self, images):
            eng = self
            assert eng.pred._pending is None    # at most one pass in flight on the encoder stream
            eng.launched += 1
            feats = self.encode_images(images)

            class Pending:
                def wait(self):
                    eng.joined += 1
                    return feats
            return Pending()

    def run(overlap):
        eng = AsyncOracle(cfg, sd, fill_holes=False)
        pred = SAM2VideoPredictor(eng, fill_hole_area=0, feature_cache_frames=cache, encoder_batch_frames=4,
                                  encoder_overlap=overlap)
        assert pred.encoder_overlap == overlap
        eng.pred = pred
        res, st = _track(pred, vid, 12, reverse=reverse, chunks=chunks)
        assert pred._upcoming is None and pred._prefetched == {} and pred._pending is None
        return res, eng

    sync, e0 = run(False)
    over, e1 = run(True)
    assert e0.launched == 0 and e1.launched > 0
    assert e1.joined == e1.launched              # a pass is only launched for frames the call will reach
    assert e1.frames_encoded == e0.frames_encoded
    _same(sync, over, exact=True)


def test_overlap_needs_an_engine_with_an_encoder_stream():
    from oracle import sam2_oracle as O
    cfg = get_config("tiny", image_size=256)
    pred = SAM2VideoPredictor(O.OracleEngine(cfg, synthetic_state_dict(cfg, 0), fill_holes=False), encoder_overlap=True)
    assert pred.encoder_overlap is False


# ------------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("model,size", [("tiny", 512), ("large", 1024)])
def test_batched_encoder_is_bit_identical(model, size):
    from detsam2_b200.engine import CudaEngine
    cfg = get_config(model, image_size=size)
    eng = CudaEngine(cfg, synthetic_state_dict(cfg, 0))
    torch.manual_seed(5)
    frames = (torch.randn(3, 3, size, size, device="cuda") * 1.5).half()
    for rep in range(4):                     # eager, eager, capture, replay
        single = [eng.encode_image(frames[i]) for i in range(3)]
        batched = eng.encode_images(frames)
        for s, b in zip(single, batched):
            for name in ("vis_f32", "vis_bf16", "feat_s0", "feat_s1"):
                x, y = getattr(s, name), getattr(b, name)
                assert x.shape == y.shape and x.dtype == y.dtype
                assert torch.equal(x, y), (rep, name)
    assert len(eng.encode_images(frames[:1])) == 1


@pytest.mark.gpu
def test_tracking_is_bit_identical_with_encode_ahead():
    from detsam2_b200.engine import CudaEngine
    cfg = get_config("tiny", image_size=512)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=3, height=512, width=512, num_frames=12, seed=6)
    res = {}
    for E, overlap in ((1, False), (4, False), (4, True)):
        pred = SAM2VideoPredictor(CudaEngine(cfg, sd), fill_hole_area=8, encoder_batch_frames=E, encoder_overlap=overlap)
        assert pred.encoder_overlap == overlap
        res[E, overlap], _ = _track(pred, vid, 12, reverse=True, chunks=3)
    _same(res[1, False], res[4, False], exact=True)
    _same(res[1, False], res[4, True], exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("host_frames", [False, True])
def test_overlapped_encoder_is_bit_identical_on_a_long_forward_run(host_frames):
    """Encoder passes on the engine's own stream while the tracker runs (engine.encode_images_async): 40 frames, forward,
    the host never synchronises inside the loop (so the two streams really overlap and the CUDA graphs of both seams are
    replayed concurrently) — every mask logit equal to the run that keeps the encoder on the tracker's stream."""
    from detsam2_b200.engine import CudaEngine
    cfg = get_config("tiny", image_size=512)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=4, height=512, width=512, num_frames=40, seed=8)
    frames = [vid.frame(t) for t in range(40)]
    res = {}
    for overlap in (False, True):
        pred = SAM2VideoPredictor(CudaEngine(cfg, sd), fill_hole_area=8, encoder_batch_frames=4, encoder_overlap=overlap)
        with torch.inference_mode():
            st = pred.init_state(frames, offload_video_to_cpu=host_frames)
            for oid, b in vid.boxes(0).items():
                pred.add_new_points_or_box(st, 0, oid, box=np.asarray(b, dtype=np.float32))
            out = [(f, m) for f, _, m in pred.propagate_in_video(st)]     # device tensors: no sync until the end
            torch.cuda.synchronize()
        res[overlap] = out
        if overlap:
            assert pred.engine._enc_stream is not None
    assert len(res[False]) == len(res[True]) == 40
    for (fa, ma), (fb, mb) in zip(res[False], res[True]):
        assert fa == fb and torch.equal(ma, mb), fa


@pytest.mark.gpu
@pytest.mark.parametrize("frames_on_device", [True, False])
def test_overlapped_encoder_is_bit_identical_in_stream_mode(frames_on_device):
    """Det-SAM2's own drive mode (VideoProcessor: chunks, detection every few frames, reverse re-tracking over a window,
    release of old frames, feature cache) with the encoder passes on the engine's stream and without: every handed-out mask
    identical.  Long enough (40 frames, K = 8, M = 16) that reverse windows span several encoder passes."""
    from detsam2_b200.engine import CudaEngine
    from detsam2_b200.synthetic import GroundTruthDetector
    from detsam2_b200.video_processor import VideoProcessor
    cfg = get_config("tiny", image_size=512)
    sd = synthetic_state_dict(cfg, 0)
    vid = BilliardVideo(num_objects=3, height=240, width=320, num_frames=40, seed=9)
    res = {}
    for overlap in (False, True):
        pred = SAM2VideoPredictor(CudaEngine(cfg, sd), fill_hole_area=8, encoder_batch_frames=4, encoder_overlap=overlap,
                                  feature_cache_frames=20)
        vp = VideoProcessor(predictor=pred, detector=GroundTruthDetector(vid, detect_interval=8), frame_buffer_size=8,
                            detect_interval=8, max_frame_num_to_track=16, max_inference_state_frames=16, skip_classes=set(),
                            frames_on_device=frames_on_device)
        with torch.inference_mode():
            res[overlap] = vp.run(frames=(vid.frame(t) for t in range(40)))
        assert pred.encoder_overlap == overlap and pred._pending is None
    assert sorted(res[False]) == sorted(res[True]) == list(range(40))
    for t in range(40):
        assert sorted(res[False][t]) == sorted(res[True][t])
        for oid, m in res[False][t].items():
            assert np.array_equal(m, res[True][t][oid]), (t, oid)
