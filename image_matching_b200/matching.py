"""Drop-in replacements for the reference's inference modules, backed by libb200match.so.

    reference                                                      here
    superglue/models/matching_test.py:47-82   Matching          -> Matching
    superpoint/models/superpoint_test.py:55-161 SuperPoint      -> SuperPoint
    superglue/models/superglue_test.py:177-285 SuperGlue        -> SuperGlue

Same constructor config dicts and defaults, same state_dict key names (so the reference's
checkpoints load with ``load_state_dict``), same ``forward`` inputs/outputs (container
types, dtypes, layouts).  torch is only the boundary: tensors in / tensors out, device
memory and the current CUDA stream.  All arithmetic happens in the hand-written sm_100a
kernels behind the C ABI; there is no eager/CPU fallback -- CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import lib as _lib


def _align_corners_from_torch_version() -> bool:
    # the reference's switch, verbatim semantics (superpoint_test.py:47): third character of the
    # version string; "2.11.0" -> '1' -> False
    try:
        return int(torch.__version__[2]) > 2
    except ValueError:
        return False


def _image(x: torch.Tensor) -> torch.Tensor:
    """(B,1,H,W) image batch as the library takes it: fp32 in [0,1] (the reference's contract), or -- extension,
    SURVEY.md 8 f2 -- the raw uint8 pixels, normalised by 255 on the device exactly like datasets/SSHIDataset.py:26-29
    followed by `.float()`, so the 8-bit image can be uploaded instead of a 4x (float64: 8x) larger tensor."""
    if x.dim() != 4 or x.shape[1] != 1:
        # the reference's first Conv2d(1, 64) raises on anything but one channel; never read a (B,C>1,H,W) buffer as B*H*W
        raise ValueError(f"expected a (B,1,H,W) grayscale image batch, got shape {tuple(x.shape)}")
    return x.contiguous() if x.dtype == torch.uint8 else x.contiguous().float()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool = False):
    """Create nested container modules so that state_dict() yields exactly `dotted`."""
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, nn.Module())
        m = m._modules[p]
    if buffer:
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _conv_bn_params(root, conv, bn, cout, cin, k, ndim=2):
    shape = (cout, cin, k, k) if ndim == 2 else (cout, cin, 1)
    fan_in = cin * k * k
    bound = 1.0 / np.sqrt(fan_in)
    _register(root, conv + ".weight", torch.empty(shape).uniform_(-bound, bound))
    _register(root, conv + ".bias", torch.empty(cout).uniform_(-bound, bound))
    if bn:
        _register(root, bn + ".weight", torch.ones(cout))
        _register(root, bn + ".bias", torch.zeros(cout))
        _register(root, bn + ".running_mean", torch.zeros(cout), buffer=True)
        _register(root, bn + ".running_var", torch.ones(cout), buffer=True)
        _register(root, bn + ".num_batches_tracked", torch.tensor(0, dtype=torch.long), buffer=True)


class _Engine:
    """Owns one b200m handle (per device) for a set of modules and their packed weights."""

    def __init__(self):
        self.handle = None
        self.device = None
        self.cfg_key = None
        self.dirty = True
        self.ws = None
        self.packed = {}        # module id -> _weights_version this engine last packed
        self._staging = {}

    def close(self):
        if self.handle is not None:
            _lib.load().b200m_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ensure(self, device: torch.device, sp: "SuperPoint | None", sg: "SuperGlue | None"):
        if device.type != "cuda":
            raise RuntimeError("image_matching_b200 runs on CUDA (sm_100a) only; got a %s tensor -- "
                               "there is no CPU fallback" % device.type)
        L = _lib.load()
        spc = sp.config if sp is not None else SuperPoint.default_config
        # a SuperPoint-only engine never runs the SuperGlue half: give it a minimal, always-valid shape
        sgc = sg.config if sg is not None else {"keypoint_encoder": [32], "GNN_layers": [],
                                                "sinkhorn_iterations": 0, "match_threshold": 0.2}
        D = (sp or sg).config["descriptor_dim"]
        names = list(sgc["GNN_layers"])
        key = (device.index, D, spc["nms_radius"], float(spc["keypoint_threshold"]), spc["max_keypoints"],
               spc["remove_borders"], tuple(sgc["keypoint_encoder"]), tuple(names),
               sgc["sinkhorn_iterations"], float(sgc["match_threshold"]))
        if self.handle is None or key != self.cfg_key:
            self.close()
            cfg = _lib.Config()
            cfg.descriptor_dim = D
            cfg.nms_radius = spc["nms_radius"]
            cfg.keypoint_threshold = spc["keypoint_threshold"]
            cfg.max_keypoints = spc["max_keypoints"]
            cfg.remove_borders = spc["remove_borders"]
            cfg.align_corners = int(_align_corners_from_torch_version())
            kenc = list(sgc["keypoint_encoder"])
            cfg.n_kenc = len(kenc)
            for i, v in enumerate(kenc):
                cfg.kenc[i] = v
            cfg.n_gnn_layers = len(names)
            for i, nme in enumerate(names):
                cfg.gnn_cross[i] = 1 if nme == "cross" else 0
            cfg.sinkhorn_iterations = sgc["sinkhorn_iterations"]
            cfg.match_threshold = sgc["match_threshold"]
            hp = C.c_void_p()
            idx = device.index if device.index is not None else torch.cuda.current_device()
            _lib.check(L.b200m_create(C.byref(cfg), idx, C.byref(hp)), "b200m_create")
            self.handle, self.device, self.cfg_key, self.dirty = hp, device, key, True
            self.packed = {}
        # every engine remembers the weight version IT packed: a module shared by two engines (its own and Matching's)
        # is repacked by each of them when its weights change
        stale = [mod for mod in (sp, sg) if mod is not None and self.packed.get(id(mod)) != mod._weights_version]
        if self.dirty or stale:
            for prefix, mod in (("superpoint.", sp), ("superglue.", sg)):
                if mod is None:
                    continue
                for k, v in mod._engine_state().items():
                    if k.endswith("num_batches_tracked"):
                        continue
                    a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
                    shp = (C.c_int64 * max(a.ndim, 1))(*a.shape)
                    _lib.check(L.b200m_set_tensor(self.handle, (prefix + k).encode(),
                                                  a.ctypes.data_as(C.c_void_p), shp, a.ndim), "b200m_set_tensor")
                self.packed[id(mod)] = mod._weights_version
            _lib.check(L.b200m_pack(self.handle, _stream(device)), "b200m_pack")
            self.dirty = False
        return L

    def workspace(self, nbytes: int, device):
        if self.ws is None or self.ws.numel() < nbytes or self.ws.device != device:
            self.ws = None
            self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.ws

    def launch_count(self) -> int:
        return int(_lib.load().b200m_launch_count(self.handle)) if self.handle is not None else 0

    def graph_replays(self) -> int:
        return int(_lib.load().b200m_graph_replay_count(self.handle)) if self.handle is not None else 0

    def staging(self, key, nbytes: int, device):
        """Persistent (per shape key) flat device buffer: stable pointers let the library replay its captured CUDA graph."""
        buf = self._staging.get(key)
        if buf is None or buf.numel() < nbytes or buf.device != device:
            if len(self._staging) > 8:
                self._staging.clear()
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            self._staging[key] = buf
        return buf


def _stream(device):
    """torch's current stream ON THE TENSORS' DEVICE (not the process's current device); the C entry points switch to
    the handle's device themselves and restore the caller's."""
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _B200Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._weights_version = 0
        self._engine = _Engine()

    def mark_dirty(self):
        """Tell every engine using this module that its weights changed.  load_state_dict() does this itself; call it
        after editing parameters in place (``p.data.copy_()``, ``bin_score.fill_()``), which torch does not report."""
        self._weights_version += 1

    def _load_from_state_dict(self, *a, **k):   # weights changed -> repack on next forward
        self._weights_version += 1
        return super()._load_from_state_dict(*a, **k)

    def load_state_dict(self, *a, **k):
        self._weights_version += 1
        return super().load_state_dict(*a, **k)

    def _engine_state(self):
        """name -> tensor as the C library expects them (b200m_set_tensor); subclasses with other key names remap."""
        return self.state_dict()


class SuperPoint(_B200Module):
    """B200-native SuperPoint (reference: superpoint/models/superpoint_test.py:55-161)."""
    default_config = {
        "descriptor_dim": 256,
        "nms_radius": 4,
        "keypoint_threshold": 0.005,
        "max_keypoints": -1,
        "remove_borders": 4,
    }

    def __init__(self, config):
        super().__init__()
        self.config = {**self.default_config, **config}
        d1 = self.config["descriptor_dim"]
        c1, c2, c3, c4, c5, det_h = 64, 64, 128, 128, 256, 65
        for prefix, cin, cout in (("inc.conv.conv", 1, c1), ("down1.mpconv.1.conv", c1, c2),
                                  ("down2.mpconv.1.conv", c2, c3), ("down3.mpconv.1.conv", c3, c4)):
            _conv_bn_params(self, prefix + ".0", prefix + ".1", cout, cin, 3)
            _conv_bn_params(self, prefix + ".3", prefix + ".4", cout, cout, 3)
        _conv_bn_params(self, "convPa", "bnPa", c5, c4, 3)
        _conv_bn_params(self, "convPb", "bnPb", det_h, c5, 1)
        _conv_bn_params(self, "convDa", "bnDa", c5, c4, 3)
        _conv_bn_params(self, "convDb", "bnDb", d1, c5, 1)
        if self.config["weights"]:   # KeyError if absent, like the reference (superpoint_test.py:87)
            checkpoints = torch.load(self.config["weights"], map_location="cpu")
            sd = OrderedDict((k[7:] if "module" in k else k, v)
                             for k, v in checkpoints["model_state_dict"].items())
            self.load_state_dict(sd)
            print("Loaded SuperPoint model")

    def _run(self, engine, L, x, sg=None):
        B, _, H, W = x.shape
        dev = x.device
        D = self.config["descriptor_dim"]
        cap = int(L.b200m_keypoint_capacity(engine.handle, H, W))
        kp = torch.empty((B, cap, 2), dtype=torch.float32, device=dev)
        sc = torch.empty((B, cap), dtype=torch.float32, device=dev)
        de = torch.empty((B, D, cap), dtype=torch.float32, device=dev)
        cnt = torch.empty((B,), dtype=torch.int32, device=dev)
        nbytes = L.b200m_superpoint_workspace_bytes(engine.handle, B, H, W)
        ws = engine.workspace(nbytes, dev)
        fn = L.b200m_superpoint_forward_u8 if x.dtype == torch.uint8 else L.b200m_superpoint_forward
        _lib.check(fn(engine.handle, _ptr(x), B, H, W, _ptr(kp), _ptr(sc), _ptr(de),
                      _ptr(cnt), cap, _ptr(ws), ws.numel(), _stream(dev)), "b200m_superpoint_forward")
        return kp, sc, de, cnt

    @staticmethod
    def _check_counts(cnt_host):
        """Negative counts are the library's sticky device-side error flags (read back with the counts)."""
        if any(c == -2 for c in cnt_host):
            raise RuntimeError("SuperPoint activation left the fp16 range of the tensor-core convolution "
                               "(|x| > 65504); set B200M_CONV_IMPL=simt for the fp32 CUDA-core path")
        if any(c == -1 for c in cnt_host):
            raise RuntimeError("keypoint candidate list overflowed (degenerate heat-map with massive ties)")

    @staticmethod
    def _to_lists(kp, sc, de, cnt_host):
        SuperPoint._check_counts(cnt_host)
        cap = kp.shape[1]
        # one unbind per tensor (views) instead of a Python-level slice per image; only short images are narrowed
        keypoints, scores, descriptors = list(kp.unbind(0)), list(sc.unbind(0)), list(de.unbind(0))
        for i, n in enumerate(cnt_host):
            if n != cap:
                keypoints[i], scores[i], descriptors[i] = keypoints[i][:n], scores[i][:n], descriptors[i][:, :n]
        return keypoints, tuple(scores), descriptors

    def forward(self, x):
        x = _image(x)
        L = self._engine.ensure(x.device, self, None)
        kp, sc, de, cnt = self._run(self._engine, L, x)
        keypoints, scores, descriptors = self._to_lists(kp, sc, de, cnt.cpu().tolist())
        return {"keypoints": keypoints, "scores": scores, "descriptors": descriptors}


def knn_ratio_match(superpoint: "SuperPoint", desc0: torch.Tensor, desc1: torch.Tensor, ratio: float = 0.7):
    """GPU replacement for the FLANN step of superpoint_flann_test.py:69-78: exact 2-nearest-neighbour search between two
    SuperPoint descriptor sets + Lowe's ratio test.  desc0 (D,N) / desc1 (D,M) (one image pair, as `pred['descriptors'][0]`)
    or batched (B,D,N) / (B,D,M).  Returns (matches, dist1, dist2): index of the nearest descriptor of desc1 or -1
    where `dist1 < ratio * dist2` fails, and the Euclidean distances to the two nearest."""
    single = desc0.dim() == 2
    d0 = (desc0[None] if single else desc0).contiguous().float()
    d1 = (desc1[None] if single else desc1).contiguous().float()
    L = superpoint._engine.ensure(d0.device, superpoint, None)
    B, D, N = d0.shape
    M = d1.shape[2]
    if D != superpoint.config["descriptor_dim"] or d1.shape[0] != B or d1.shape[1] != D:
        raise ValueError("descriptor shapes do not match the model")
    match = torch.empty((B, N), dtype=torch.int64, device=d0.device)
    e1 = torch.empty((B, N), dtype=torch.float32, device=d0.device)
    e2 = torch.empty((B, N), dtype=torch.float32, device=d0.device)
    _lib.check(L.b200m_knn_ratio_match(superpoint._engine.handle, _ptr(d0), _ptr(d1), None, None, B, N, M,
                                       float(ratio), _ptr(match), _ptr(e1), _ptr(e2), _stream(d0.device)), "b200m_knn_ratio_match")
    return (match[0], e1[0], e2[0]) if single else (match, e1, e2)


class SuperPointOfficial(SuperPoint):
    """The "official" (MagicLeap) SuperPoint of the reference, superglue/models/superpoint.py:95-202: the same
    encoder / heads topology WITHOUT BatchNorm and with the key names ``conv1a .. conv4b, convPa/Pb/Da/Db`` (so the
    public ``superpoint_v1.pth`` loads with ``load_state_dict``).  Same kernels: the library uses a convolution as is
    when no BatchNorm tensors accompany it.  Differences to the reference kept: ``max_keypoints`` validation (:143-145);
    the dense map is normalised with an eps-clamped norm (:188), identical for any non-zero descriptor."""
    _NAMES = OrderedDict([("conv1a", "inc.conv.conv.0"), ("conv1b", "inc.conv.conv.3"),
                          ("conv2a", "down1.mpconv.1.conv.0"), ("conv2b", "down1.mpconv.1.conv.3"),
                          ("conv3a", "down2.mpconv.1.conv.0"), ("conv3b", "down2.mpconv.1.conv.3"),
                          ("conv4a", "down3.mpconv.1.conv.0"), ("conv4b", "down3.mpconv.1.conv.3"),
                          ("convPa", "convPa"), ("convPb", "convPb"), ("convDa", "convDa"), ("convDb", "convDb")])

    def __init__(self, config):
        _B200Module.__init__(self)
        self.config = {**self.default_config, **config}
        c1, c2, c3, c4, c5 = 64, 64, 128, 128, 256
        shapes = {"conv1a": (c1, 1, 3), "conv1b": (c1, c1, 3), "conv2a": (c2, c1, 3), "conv2b": (c2, c2, 3),
                  "conv3a": (c3, c2, 3), "conv3b": (c3, c3, 3), "conv4a": (c4, c3, 3), "conv4b": (c4, c4, 3),
                  "convPa": (c5, c4, 3), "convPb": (65, c5, 1), "convDa": (c5, c4, 3),
                  "convDb": (self.config["descriptor_dim"], c5, 1)}
        for name, (cout, cin, k) in shapes.items():
            _conv_bn_params(self, name, None, cout, cin, k)
        # the reference loads <its dir>/weights/superpoint_v1.pth unconditionally (:140-141; a git-LFS stub in the
        # reference tree); here the path comes from config['weights'] (as superpoint_glue_official_test.py:40 passes
        # it) and a falsy / missing entry leaves the seeded initialisation for load_state_dict()
        path = self.config.get("weights")
        if path:
            self.load_state_dict(torch.load(str(path), map_location="cpu"))
            print("Loaded SuperPoint model")
        mk = self.config["max_keypoints"]
        if mk == 0 or mk < -1:
            raise ValueError('"max_keypoints" must be positive or "-1"')

    def _engine_state(self):
        sd = self.state_dict()
        return OrderedDict((self._NAMES[k.rsplit(".", 1)[0]] + "." + k.rsplit(".", 1)[1], v) for k, v in sd.items())

    def forward(self, data):
        """reference signature: data = {'image': (B,1,H,W)} (superpoint.py:149)."""
        return SuperPoint.forward(self, data["image"] if isinstance(data, dict) else data)


class SuperGlue(_B200Module):
    """B200-native SuperGlue (reference: superglue/models/superglue_test.py:177-285)."""
    default_config = {
        "descriptor_dim": 256,
        "weights": "indoor",
        "keypoint_encoder": [32, 64, 128, 256],
        "GNN_layers": ["self", "cross"] * 9,
        "sinkhorn_iterations": 100,
        "match_threshold": 0.2,
    }

    def __init__(self, config):
        super().__init__()
        self.config = {**self.default_config, **config}
        D = self.config["descriptor_dim"]
        ch = [3] + list(self.config["keypoint_encoder"]) + [D]
        idx = 0
        for i in range(1, len(ch)):
            last = i == len(ch) - 1
            _conv_bn_params(self, f"kenc.encoder.{idx}", None if last else f"kenc.encoder.{idx + 1}",
                            ch[i], ch[i - 1], 1, ndim=1)
            idx += 1 if last else 3
        for l in range(len(self.config["GNN_layers"])):
            p = f"gnn.layers.{l}"
            _conv_bn_params(self, p + ".attn.merge", None, D, D, 1, ndim=1)
            for j in range(3):
                _conv_bn_params(self, p + f".attn.proj.{j}", None, D, D, 1, ndim=1)
            _conv_bn_params(self, p + ".mlp.0", p + ".mlp.1", 2 * D, 2 * D, 1, ndim=1)
            _conv_bn_params(self, p + ".mlp.3", None, D, 2 * D, 1, ndim=1)
        _conv_bn_params(self, "final_proj", None, D, D, 1, ndim=1)
        _register(self, "bin_score", torch.tensor(1.0))
        # state_dict order of the reference: bin_score first; order does not matter for loading
        if self.config["weights"]:
            checkpoints = torch.load(config["weights"], map_location="cpu")
            if "indoor" in self.config["weights"] or "outdoor" in self.config["weights"]:
                state_dict = checkpoints
            else:
                state_dict = checkpoints["net"]
            self.load_state_dict(state_dict)
            print("Loaded SuperGlue model weights")

    def _run(self, engine, L, data, counts0=None, counts1=None):
        kpts0, kpts1 = data["keypoints0"].contiguous().float(), data["keypoints1"].contiguous().float()
        desc0, desc1 = data["descriptors0"].contiguous().float(), data["descriptors1"].contiguous().float()
        sc0, sc1 = data["scores0"].contiguous().float(), data["scores1"].contiguous().float()
        B, N, M = kpts0.shape[0], kpts0.shape[1], kpts1.shape[1]
        dev = kpts0.device
        _, _, H0, W0 = data["image0"].shape
        _, _, H1, W1 = data["image1"].shape
        m0 = torch.empty((B, N), dtype=torch.int64, device=dev)
        m1 = torch.empty((B, M), dtype=torch.int64, device=dev)
        s0 = torch.empty((B, N), dtype=torch.float32, device=dev)
        s1 = torch.empty((B, M), dtype=torch.float32, device=dev)
        nbytes = L.b200m_superglue_workspace_bytes(engine.handle, B, N, M)
        ws = engine.workspace(nbytes, dev)
        _lib.check(L.b200m_superglue_forward(engine.handle, _ptr(kpts0), _ptr(sc0), _ptr(desc0), _ptr(counts0),
                                             _ptr(kpts1), _ptr(sc1), _ptr(desc1), _ptr(counts1),
                                             B, N, M, H0, W0, H1, W1, _ptr(m0), _ptr(m1), _ptr(s0), _ptr(s1),
                                             _ptr(ws), ws.numel(), _stream(dev)), "b200m_superglue_forward")
        return m0, m1, s0, s1

    def forward(self, data, _engine=None):
        kpts0, kpts1 = data["keypoints0"], data["keypoints1"]
        if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:  # no keypoints (superglue_test.py:235-242)
            shape0, shape1 = kpts0.shape[:-1], kpts1.shape[:-1]
            return {
                "matches0": kpts0.new_full(shape0, -1, dtype=torch.int),
                "matches1": kpts1.new_full(shape1, -1, dtype=torch.int),
                "matching_scores0": kpts0.new_zeros(shape0),
                "matching_scores1": kpts1.new_zeros(shape1),
            }
        engine = _engine or self._engine
        L = engine.ensure(kpts0.device, None if _engine is None else _engine._sp, self)
        m0, m1, s0, s1 = self._run(engine, L, data)
        return {"matches0": m0, "matches1": m1, "matching_scores0": s0, "matching_scores1": s1}


_OUT_LAYOUT = (("keypoints0", "f", lambda B, c, D: (B, c, 2)), ("scores0", "f", lambda B, c, D: (B, c)),
               ("descriptors0", "f", lambda B, c, D: (B, D, c)), ("keypoints1", "f", lambda B, c, D: (B, c, 2)),
               ("scores1", "f", lambda B, c, D: (B, c)), ("descriptors1", "f", lambda B, c, D: (B, D, c)),
               ("counts", "i", lambda B, c, D: (2, B)), ("matches0", "l", lambda B, c, D: (B, c)),
               ("matches1", "l", lambda B, c, D: (B, c)), ("matching_scores0", "f", lambda B, c, D: (B, c)),
               ("matching_scores1", "f", lambda B, c, D: (B, c)))
_OUT_DTYPES = {"f": (torch.float32, 4), "i": (torch.int32, 4), "l": (torch.int64, 8)}


def _out_bytes(B, cap, D):
    n = 0
    for _, t, shp in _OUT_LAYOUT:
        n = (n + 255) // 256 * 256 + int(np.prod(shp(B, cap, D))) * _OUT_DTYPES[t][1]
    return n + 256


def _out_views(flat: torch.Tensor, B, cap, D):
    """The forward_device result dict as views into one flat uint8 buffer (256-byte aligned sub-tensors)."""
    out, off = {}, 0
    assert flat.data_ptr() % 256 == 0           # torch's CUDA allocations are 512-byte aligned
    for name, t, shp in _OUT_LAYOUT:
        dt, sz = _OUT_DTYPES[t]
        shape = shp(B, cap, D)
        off = (off + 255) // 256 * 256
        nb = int(np.prod(shape)) * sz
        out[name] = flat[off: off + nb].view(dt).view(shape)
        off += nb
    return out


class Matching(nn.Module):
    """Image Matching Frontend (SuperPoint + SuperGlue); reference: superglue/models/matching_test.py:47-82."""

    _superpoint_cls = None   # set below (SuperPoint); MatchingOfficial swaps in SuperPointOfficial

    def __init__(self, config={}):
        super().__init__()
        self.superpoint = (self._superpoint_cls or SuperPoint)(config.get("superpoint", {}))
        self.superglue = SuperGlue(config.get("superglue", {}))
        self._engine = _Engine()
        self._engine._sp = self.superpoint

    def _ensure(self, device):
        return self._engine.ensure(device, self.superpoint, self.superglue)

    # ---- fused fast path: one C call for SuperPoint x2 + SuperGlue, no host sync inside
    def forward_device(self, image0: torch.Tensor, image1: torch.Tensor, out: dict | None = None):
        """Device-resident results without any host synchronisation: dict of padded tensors plus
        per-pair `counts0/1` (number of valid leading keypoints).  `out`: a dict returned by an earlier call with the
        same shapes, to be overwritten in place -- with unchanged input, output and workspace pointers the library
        replays the CUDA graph it captured for this call instead of re-launching ~100 kernels one by one."""
        image0, image1 = _image(image0), _image(image1)
        if image0.dtype != image1.dtype:
            image0, image1 = image0.float() / (255.0 if image0.dtype == torch.uint8 else 1.0), \
                image1.float() / (255.0 if image1.dtype == torch.uint8 else 1.0)
        if image0.shape != image1.shape:
            raise ValueError("forward_device needs equally shaped image batches")
        L = self._ensure(image0.device)
        e = self._engine
        B, _, H, W = image0.shape
        dev = image0.device
        D = self.superpoint.config["descriptor_dim"]
        if self.superpoint.config["max_keypoints"] < 0:
            return self._forward_device_unbounded(L, image0, image1)
        cap = int(L.b200m_keypoint_capacity(e.handle, H, W))
        if out is None or out["keypoints0"].shape != (B, cap, 2) or out["keypoints0"].device != dev:
            out = _out_views(torch.empty(_out_bytes(B, cap, D), dtype=torch.uint8, device=dev), B, cap, D)
        nbytes = L.b200m_matching_workspace_bytes(e.handle, B, H, W)
        ws = e.workspace(nbytes, dev)
        c0, c1 = out["counts"][0], out["counts"][1]
        fn = L.b200m_matching_forward_u8 if image0.dtype == torch.uint8 else L.b200m_matching_forward
        _lib.check(fn(
            e.handle, _ptr(image0), _ptr(image1), B, H, W,
            _ptr(out["keypoints0"]), _ptr(out["scores0"]), _ptr(out["descriptors0"]), _ptr(c0),
            _ptr(out["keypoints1"]), _ptr(out["scores1"]), _ptr(out["descriptors1"]), _ptr(c1), cap,
            _ptr(out["matches0"]), _ptr(out["matches1"]), _ptr(out["matching_scores0"]),
            _ptr(out["matching_scores1"]), _ptr(ws), ws.numel(), _stream(dev)), "b200m_matching_forward")
        return out

    def _forward_replayable(self, image0, image1):
        """forward_device for the reference-facing ``forward``: the caller hands in fresh tensors on every call and must
        get fresh tensors back, but a CUDA graph replays only with unchanged pointers.  So (for batches up to 256 MB of
        pixels) the images are copied into a persistent staging buffer, the library writes into a persistent result
        buffer, and ONE device copy of that buffer becomes the caller's result -- three small copies buy the replay of
        ~100 launches, which is what bounds the reference's own batch_size = 1 loop (superpoint_glue_test.py:65-78)."""
        image0, image1 = _image(image0), _image(image1)
        nbytes = image0.numel() * image0.element_size()
        if (self.superpoint.config["max_keypoints"] < 0 or image0.dtype != image1.dtype or image0.shape != image1.shape
                or nbytes > (128 << 20)):
            return self.forward_device(image0, image1)
        L = self._ensure(image0.device)
        e = self._engine
        B, _, H, W = image0.shape
        D = self.superpoint.config["descriptor_dim"]
        cap = int(L.b200m_keypoint_capacity(e.handle, H, W))
        key = (B, H, W, cap, D, image0.dtype)
        nb_al = (nbytes + 255) // 256 * 256
        stage = e.staging(("in",) + key, 2 * nb_al, image0.device)
        s0 = stage[:nbytes].view(image0.dtype).view(image0.shape)
        s1 = stage[nb_al:nb_al + nbytes].view(image0.dtype).view(image0.shape)
        s0.copy_(image0)
        s1.copy_(image1)
        flat = e.staging(("out",) + key, _out_bytes(B, cap, D), image0.device)
        self.forward_device(s0, s1, out=_out_views(flat, B, cap, D))
        return _out_views(flat.clone(), B, cap, D)

    def _forward_device_unbounded(self, L, image0, image1):
        """``max_keypoints = -1`` (the default of SuperPoint.default_config and of superpoint_glue_test.py:29): the
        keypoint count is only bounded by the NMS capacity (16384 at 640x480, 65536 at 1280x960), and padding SuperGlue
        to that bound would cost cap^2 score matrices (1 GiB / 17 GiB per pair) and 8-16x redundant GNN work.  So this
        configuration runs in two phases: SuperPoint, ONE host read of the per-image counts, then SuperGlue sized to the
        largest count actually found.  Same output dict as forward_device, padded to that count."""
        e = self._engine
        B = image0.shape[0]
        dev = image0.device
        kp0, sc0, de0, c0 = self.superpoint._run(e, L, image0)
        kp1, sc1, de1, c1 = self.superpoint._run(e, L, image1)
        counts = torch.stack([c0, c1])
        nmax = counts.max(dim=1).values.cpu().tolist()          # the host sync of this configuration
        n, m = max(int(nmax[0]), 0), max(int(nmax[1]), 0)
        out = {"keypoints0": kp0[:, :n].contiguous(), "scores0": sc0[:, :n].contiguous(),
               "descriptors0": de0[:, :, :n].contiguous(),
               "keypoints1": kp1[:, :m].contiguous(), "scores1": sc1[:, :m].contiguous(),
               "descriptors1": de1[:, :, :m].contiguous(), "counts": counts}
        if n == 0 or m == 0:
            out.update({"matches0": torch.full((B, n), -1, dtype=torch.int64, device=dev),
                        "matches1": torch.full((B, m), -1, dtype=torch.int64, device=dev),
                        "matching_scores0": torch.zeros((B, n), device=dev),
                        "matching_scores1": torch.zeros((B, m), device=dev)})
            return out
        data = {"image0": image0, "image1": image1, "keypoints0": out["keypoints0"], "scores0": out["scores0"],
                "descriptors0": out["descriptors0"], "keypoints1": out["keypoints1"], "scores1": out["scores1"],
                "descriptors1": out["descriptors1"]}
        m0, m1, s0, s1 = self.superglue._run(e, L, data, counts0=c0, counts1=c1)
        out.update({"matches0": m0, "matches1": m1, "matching_scores0": s0, "matching_scores1": s1})
        return out

    def forward(self, data):
        """Run SuperPoint (optionally) and SuperGlue; same contract as the reference's Matching.forward."""
        pred = {}
        need0, need1 = "keypoints0" not in data, "keypoints1" not in data
        if need0 and need1 and data["image0"].shape == data["image1"].shape:
            out = self._forward_replayable(data["image0"], data["image1"])
            cnt = out["counts"].cpu()          # the only host sync: list lengths are data dependent
            c0, c1 = cnt[0].tolist(), cnt[1].tolist()
            k0, s0, d0 = SuperPoint._to_lists(out["keypoints0"], out["scores0"], out["descriptors0"], c0)
            k1, s1, d1 = SuperPoint._to_lists(out["keypoints1"], out["scores1"], out["descriptors1"], c1)
            pred = {"keypoints0": k0, "scores0": s0, "descriptors0": d0,
                    "keypoints1": k1, "scores1": s1, "descriptors1": d1}
            if len(set(c0)) > 1 or len(set(c1)) > 1:
                # the reference stacks the per-image lists (matching_test.py:75-77) and fails here
                raise RuntimeError("stack expects each tensor to be equal size, but got keypoint counts "
                                   f"{c0} / {c1} in the batch")
            n, m = c0[0], c1[0]
            if n == 0 or m == 0:
                dev = data["image0"].device
                B = len(c0)
                pred.update({"matches0": torch.full((B, n), -1, dtype=torch.int, device=dev),
                             "matches1": torch.full((B, m), -1, dtype=torch.int, device=dev),
                             "matching_scores0": torch.zeros((B, n), device=dev),
                             "matching_scores1": torch.zeros((B, m), device=dev)})
            else:
                pred.update({"matches0": out["matches0"][:, :n], "matches1": out["matches1"][:, :m],
                             "matching_scores0": out["matching_scores0"][:, :n],
                             "matching_scores1": out["matching_scores1"][:, :m]})
            return pred

        # general path (features supplied for one or both sides; matching_test.py:63-80)
        dev = data["image0"].device
        L = self._ensure(dev)
        for side, need in (("0", need0), ("1", need1)):
            if need:
                x = _image(data["image" + side])
                kp, sc, de, cnt = self.superpoint._run(self._engine, L, x)
                ks, ss, ds = SuperPoint._to_lists(kp, sc, de, cnt.cpu().tolist())
                pred.update({"keypoints" + side: ks, "scores" + side: ss, "descriptors" + side: ds})
        data = {**data, **pred}
        for k in data:
            if isinstance(data[k], (list, tuple)):
                data[k] = torch.stack(data[k])
        pred = {**pred, **self.superglue.forward(data, _engine=self._engine)}
        return pred


class MatchingOfficial(Matching):
    """reference: superglue/models/matching.py:46-82 -- the official SuperPoint (no BatchNorm) + the same SuperGlue."""
    _superpoint_cls = SuperPointOfficial
