"""Integer / bookkeeping half of the memory bank: which stored frames one tracking step may attend to.

Mirrors the selection logic at the top of ``SAM2Base._prepare_memory_conditioned_features``
(/root/reference/sam2/modeling/sam2_base.py:513-562, 588-621) and ``select_closest_cond_frames``
(/root/reference/sam2/modeling/sam2_utils.py:19-66, including Det-SAM2's rule that every preload
conditioning frame is force-included once the cap of 20 is exceeded).  Pure Python, no tensors are
touched: the result is a *plan* — an ordered list of per-frame blocks — that the CUDA engine turns
into gather launches (paged-KV style), so the bank is never concatenated on the host.
"""
from dataclasses import dataclass, field
from typing import Any, List, Tuple


def select_closest_cond_frames(frame_idx, cond_frame_outputs, max_cond_frame_num, preload_idx=None):
    """Returns (selected, unselected) dicts; dict order matters (it is the memory row order)."""
    if max_cond_frame_num == -1 or len(cond_frame_outputs) <= max_cond_frame_num:
        return cond_frame_outputs, {}
    assert max_cond_frame_num >= 2, "we should allow using 2+ conditioning frames"
    selected = {}
    before = max((t for t in cond_frame_outputs if t < frame_idx), default=None)
    if before is not None:
        selected[before] = cond_frame_outputs[before]
    after = min((t for t in cond_frame_outputs if t >= frame_idx), default=None)
    if after is not None:
        selected[after] = cond_frame_outputs[after]
    remain = max_cond_frame_num - len(selected)
    others = sorted((t for t in cond_frame_outputs if t not in selected), key=lambda t: abs(t - frame_idx))[:remain]
    selected.update((t, cond_frame_outputs[t]) for t in others)
    if preload_idx is not None:
        for t in preload_idx:
            if t not in selected:
                selected[t] = cond_frame_outputs[t]
    unselected = {t: v for t, v in cond_frame_outputs.items() if t not in selected}
    return selected, unselected


@dataclass
class MemoryPlan:
    # (index into maskmem_tpos_enc, stored frame output) in memory-row order
    frames: List[Tuple[int, Any]] = field(default_factory=list)
    # (signed temporal distance, stored frame output) for object-pointer tokens, in row order
    ptrs: List[Tuple[int, Any]] = field(default_factory=list)
    t_diff_max: int = 1

    def num_tokens(self, tokens_per_frame, tokens_per_ptr=4):
        return len(self.frames) * tokens_per_frame + len(self.ptrs) * tokens_per_ptr


def plan_memory(frame_idx, output_dict, num_frames, reverse, preload_idx, num_maskmem=7,
                max_cond_frames_in_attn=20, max_obj_ptrs_in_encoder=16, stride=1):
    cond = output_dict["cond_frame_outputs"]
    non_cond = output_dict["non_cond_frame_outputs"]
    assert len(cond) > 0
    selected, unselected = select_closest_cond_frames(frame_idx, cond, max_cond_frames_in_attn, preload_idx)
    plan = MemoryPlan()
    # conditioning frames: t_pos = 0 -> temporal embedding index num_maskmem - 1
    for out in selected.values():
        plan.frames.append((num_maskmem - 1, out))
    # up to num_maskmem - 1 neighbouring non-conditioning frames, oldest first
    for t_pos in range(1, num_maskmem):
        t_rel = num_maskmem - t_pos
        if t_rel == 1:
            prev = frame_idx + t_rel if reverse else frame_idx - t_rel
        elif not reverse:
            prev = ((frame_idx - 2) // stride) * stride - (t_rel - 2) * stride
        else:
            prev = -(-(frame_idx + 2) // stride) * stride + (t_rel - 2) * stride
        out = non_cond.get(prev, None)
        if out is None:
            out = unselected.get(prev, None)
        if out is not None:
            plan.frames.append((num_maskmem - t_pos - 1, out))
    # object pointers: conditioning frames in the past (w.r.t. tracking direction), then neighbours
    max_ptrs = min(num_frames, max_obj_ptrs_in_encoder)
    plan.t_diff_max = max_ptrs - 1
    sign = -1 if reverse else 1
    for t, out in selected.items():
        if (t >= frame_idx) if reverse else (t <= frame_idx):
            plan.ptrs.append(((frame_idx - t) * sign, out))
    for t_diff in range(1, max_ptrs):
        t = frame_idx + t_diff if reverse else frame_idx - t_diff
        if t < 0 or (num_frames is not None and t >= num_frames):
            break
        out = non_cond.get(t, unselected.get(t, None))
        if out is not None:
            plan.ptrs.append((t_diff, out))
    return plan
