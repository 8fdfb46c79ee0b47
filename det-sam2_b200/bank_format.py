"""Compact, versioned on-disk format for the preload memory bank.

The reference pickles the whole ``inference_state`` (det_sam2_RT.py:489-503).  Plain ``pickle`` writes every
tensor with its full underlying storage, and the state holds each stored frame twice — once batched in
``output_dict`` and once per object in ``output_dict_per_obj`` as views of the same memory (svp:1027-1058) — so the
file is several times larger than the data (238 MB for 12 tiny-model frames in SURVEY's probe).  This format keeps
the state's exact structure and values but

  * writes every distinct tensor once and records per-object entries as slices of the batched tensor they alias
    (the aliasing is restored on load, as the reference's code relies on it),
  * stores tensors as compressed numpy arrays inside one zip (bf16 as its 16-bit pattern),
  * carries a version tag and refuses files it does not understand.

It is lossless: ``load_bank(save_bank(state))`` equals the pickle round trip value for value
(tests/test_bank_format.py).  ``VideoProcessor`` picks it for paths ending in ``.ds2bank`` and keeps pickle for
everything else, so banks remain exchangeable with the reference.
"""
import io
import json
import zipfile
from collections import OrderedDict

import numpy as np
import torch

MAGIC = "ds2bank/1"

_NP_DTYPES = {torch.float32: "f4", torch.float16: "f2", torch.float64: "f8", torch.int32: "i4", torch.int64: "i8",
              torch.int16: "i2", torch.uint8: "u1", torch.bool: "b1", torch.int8: "i1"}
_TORCH_DTYPES = {str(k): k for k in list(_NP_DTYPES) + [torch.bfloat16]}


class BankFormatError(ValueError):
    pass


def _storage_key(t):
    return (t.device.type, t.device.index, t.untyped_storage().data_ptr())


def _is_expanded(t):
    return t.dim() > 0 and t.shape[0] > 1 and t.stride(0) == 0


def _layout_perm(t):
    """Dim order of decreasing stride, or None when t is already laid out that way.  Writing ``t.permute(perm)`` keeps
    the tensor's MEMORY order on disk (the CUDA engine stores ``maskmem_features`` token-major and exposes them as an
    NCHW view, engine.encode_memory); load_bank applies the inverse permutation, so a loaded bank has the strides it
    was saved with and the engine reads it without a transposing copy."""
    if t.dim() < 2:
        return None
    perm = sorted(range(t.dim()), key=lambda d: (-t.stride(d), d))
    return None if perm == list(range(t.dim())) else perm


class _Writer:
    def __init__(self):
        self.arrays = OrderedDict()      # name -> np.ndarray
        self.by_storage = {}             # storage key -> list of (tensor, array name)
        self.tensors = []                # every tensor met, in walk order
        self.perms = {}                  # array name -> dim permutation it was written in (memory order)

    def collect(self, obj):
        if isinstance(obj, torch.Tensor):
            self.tensors.append(obj)
        elif isinstance(obj, dict):
            for v in obj.values():
                self.collect(v)
        elif isinstance(obj, (list, tuple, set, frozenset)):
            for v in obj:
                self.collect(v)

    def plan(self):
        """Largest tensor of every storage first: smaller tensors of the same storage become slices of it.  A tensor
        whose leading stride is 0 (``x.expand(B, ...)``: every stored frame's ``maskmem_pos_enc`` is such a view of the
        one copy in ``constants``, svp:1406-1435) is written as its single row and re-expanded on load."""
        groups = {}
        for t in self.tensors:
            if t.numel() > 0:
                groups.setdefault(_storage_key(t), []).append(t)
        self.bases = {}                  # storage key -> list of (base tensor, name)
        for key, ts in groups.items():
            ts = sorted(ts, key=lambda t: -(t[0:1].numel() if _is_expanded(t) else t.numel()))
            self.bases[key] = []
            for t in ts:
                if self._as_slice(t) is None:
                    b = t[0:1] if _is_expanded(t) else t
                    name = f"t{len(self.arrays)}"
                    perm = _layout_perm(b)
                    self.arrays[name] = self._to_numpy(b if perm is None else b.permute(perm))
                    if perm is not None:
                        self.perms[name] = perm
                    self.bases[key].append((b, name))

    def _as_slice(self, t):
        """(base name, start, length, expand) if t == base[start:start+length] along dim 0 (same per-row strides);
        expand = B > 0 when t is that single row broadcast B times (leading stride 0)."""
        exp = _is_expanded(t)
        for b, name in self.bases.get(_storage_key(t), []):
            if b is t:
                return (name, 0, b.shape[0], 0) if b.dim() > 0 else (name, -1, 0, 0)
            if b.dim() == 0 or t.dim() != b.dim() or t.dtype != b.dtype or t.shape[1:] != b.shape[1:]:
                continue
            if t.stride()[1:] != b.stride()[1:]:
                continue
            one_row = exp or t.shape[0] == 1     # the leading stride of a single row carries no information
            if not one_row and t.stride(0) != b.stride(0):
                continue
            off = t.storage_offset() - b.storage_offset()
            if b.shape[0] == 1:
                if off == 0 and one_row:
                    return (name, 0, 1, t.shape[0] if exp else 0)
                continue
            if b.stride(0) > 0 and off >= 0 and off % b.stride(0) == 0:
                start = off // b.stride(0)
                n = 1 if exp else t.shape[0]
                if start + n <= b.shape[0]:
                    return (name, start, n, t.shape[0] if exp else 0)
        return None

    @staticmethod
    def _to_numpy(t):
        t = t.detach().to("cpu").contiguous()
        if t.dtype == torch.bfloat16:
            return t.view(torch.int16).numpy()
        if t.dtype not in _NP_DTYPES:
            raise BankFormatError(f"unsupported tensor dtype {t.dtype}")
        return t.numpy()

    def encode(self, obj):
        if isinstance(obj, torch.Tensor):
            meta = {"__t": "tensor", "dtype": str(obj.dtype), "device": obj.device.type, "shape": list(obj.shape)}
            if obj.numel() == 0:
                return meta
            ref = self._as_slice(obj)
            if ref is None:
                raise BankFormatError("internal error: tensor without a base")  # plan() covers every tensor
            meta["base"], meta["start"], meta["len"], expand = ref
            if expand:
                meta["expand"] = expand
            return meta
        if isinstance(obj, torch.device):
            return {"__t": "device", "type": obj.type}
        if isinstance(obj, OrderedDict):
            return {"__t": "odict", "items": [[self.encode(k), self.encode(v)] for k, v in obj.items()]}
        if isinstance(obj, dict):
            return {"__t": "dict", "items": [[self.encode(k), self.encode(v)] for k, v in obj.items()]}
        if isinstance(obj, list):
            return {"__t": "list", "items": [self.encode(v) for v in obj]}
        if isinstance(obj, tuple):
            return {"__t": "tuple", "items": [self.encode(v) for v in obj]}
        if isinstance(obj, (set, frozenset)):
            return {"__t": "set", "items": [self.encode(v) for v in sorted(obj)]}
        if isinstance(obj, (np.integer,)):
            return int(obj)
        if isinstance(obj, (np.floating,)):
            return float(obj)
        if obj is None or isinstance(obj, (bool, int, float, str)):
            return obj
        raise BankFormatError(f"cannot store an object of type {type(obj).__name__} in a bank")


def save_bank(inference_state, path):
    """Writes `inference_state` (the dict of SAM2VideoPredictor.init_state, any storage device) to `path`."""
    st = dict(inference_state)
    st["cached_features"] = {}           # engine handles; recomputed on demand, as after unpickling a reference bank
    w = _Writer()
    w.collect(st)
    w.plan()
    meta = {"magic": MAGIC, "state": w.encode(st), "perms": w.perms}
    with zipfile.ZipFile(path, "w", compression=zipfile.ZIP_DEFLATED, compresslevel=4) as z:
        z.writestr("meta.json", json.dumps(meta))
        for name, arr in w.arrays.items():
            buf = io.BytesIO()
            np.save(buf, arr, allow_pickle=False)
            z.writestr(name + ".npy", buf.getvalue())
    return path


def load_bank(path, map_location=None):
    """Reads a bank written by save_bank.  Tensors go to `map_location` if given, else to the device TYPE they were
    saved from (cuda tensors to the current CUDA device, cpu when CUDA is unavailable)."""
    with zipfile.ZipFile(path, "r") as z:
        try:
            meta = json.loads(z.read("meta.json"))
        except KeyError:
            raise BankFormatError(f"{path} is not a ds2 bank (no meta.json)") from None
        if meta.get("magic") != MAGIC:
            raise BankFormatError(f"{path}: unknown bank format {meta.get('magic')!r}, this build reads {MAGIC!r}")
        cache = {}
        perms = meta.get("perms", {})

        def base(name, dtype, device):
            key = (name, device)
            if key not in cache:
                arr = np.load(io.BytesIO(z.read(name + ".npy")), allow_pickle=False)
                t = torch.from_numpy(arr)
                if dtype == torch.bfloat16:
                    t = t.view(torch.bfloat16)
                t = t.to(device)
                if name in perms:       # written in memory order: undo the permutation (a view, strides restored)
                    inv = [0] * len(perms[name])
                    for i, d in enumerate(perms[name]):
                        inv[d] = i
                    t = t.permute(inv)
                cache[key] = t
            return cache[key]

        def device_of(kind):
            if map_location is not None:
                return torch.device(map_location)
            if kind == "cuda" and torch.cuda.is_available():
                return torch.device("cuda", torch.cuda.current_device())
            return torch.device("cpu")

        def decode(o):
            if not isinstance(o, dict):
                return o
            k = o["__t"]
            if k == "tensor":
                dtype, dev = _TORCH_DTYPES[o["dtype"]], device_of(o["device"])
                if "base" not in o:
                    return torch.empty(o["shape"], dtype=dtype, device=dev)
                b = base(o["base"], dtype, dev)
                if o["start"] < 0:
                    return b
                t = b[o["start"]:o["start"] + o["len"]]
                if o.get("expand"):
                    t = t.expand(o["expand"], *t.shape[1:])
                return t
            if k == "device":
                return device_of(o["type"])
            if k == "odict":
                return OrderedDict((decode(a), decode(b)) for a, b in o["items"])
            if k == "dict":
                return {decode(a): decode(b) for a, b in o["items"]}
            if k == "list":
                return [decode(v) for v in o["items"]]
            if k == "tuple":
                return tuple(decode(v) for v in o["items"])
            if k == "set":
                return set(decode(v) for v in o["items"])
            raise BankFormatError(f"unknown node type {k!r}")

        return decode(meta["state"])
