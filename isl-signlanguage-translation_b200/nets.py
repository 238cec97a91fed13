"""The three OpenPose networks as launch plans over NHWC bf16 buffers (reference: src/model.py).

`PoseNet(kind, flat_weights)` is what `Body.model` / `Hand.model` hold. It keeps the reference module's calling
convention - `model(data: float32 [N,3,h,w] cuda) -> (PAF, heat)` or `-> heat` as float32 NCHW tensors
(model.py:207,329,407) - so the frame-level wrappers that take `body_estimation.model`
(ISL_extract_features_videos.py:49-52, ISL_Model_parameter.py:44-47) keep working, but underneath every layer is
one tcgen05 implicit-GEMM launch recorded in a C-side plan (include/islpose.h) and replayed per call.

Layout decisions (DESIGN.md has the full table):
  * conv1_1 (Cin=3) runs as a 1x1 GEMM over 3x3x3 patches gathered to 32 bf16 per pixel (K = 27 -> 32);
  * torch.cat never happens: a stage input is one buffer [feat 128 | branch outputs, each padded to 8] and the
    producers write their channel slices; weights are re-indexed to that channel order when they are packed;
  * body25's dense blocks (model.py:171-177) write their three 3x3 outputs side by side in one 3w-wide buffer;
  * the stage-output layers write bf16 into the next stage's input and/or float32 planar NCHW network outputs.
"""
import ctypes as C

import torch

from . import _lib

RELU, NONE, PRELU = "relu", "none", "prelu"

_VGG = [("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool", ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool",
        ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256), "pool",
        ("conv4_1", 256, 512), ("conv4_2", 512, 512)]

OUTPUT_CHANNELS = {"coco": (38, 19), "body25": (52, 26), "hand": (22,)}


def _up8(c):
    return (c + 7) // 8 * 8


class _Program:
    """Shape-independent description of one network: buffers (channels, pyramid level) and steps."""

    def __init__(self, kind):
        self.kind = kind
        self.bufs = {}    # name -> (channels, level)
        self.steps = []   # ("im2col", dst) | ("pool", src, dst) | ("conv", {...})
        self.outputs = []  # float32 outputs: (name, channels)

    def buf(self, name, channels, level):
        prev = self.bufs.get(name)
        if prev is not None and prev != (channels, level):
            raise ValueError("buffer %s redefined" % name)
        self.bufs[name] = (channels, level)
        return name

    def conv(self, layer, src, dst=None, cout=128, k=3, act=RELU, prelu=None, chan_map=None, f32=None, first=False):
        """src = (buffer, channel offset, channels); dst = (buffer, channel offset) | None;
        f32 = index into self.outputs | None; chan_map[i] = reference input channel read at slice channel i
        (None = zero weight)."""
        self.steps.append(("conv", dict(layer=layer, src=src, dst=dst, cout=cout, k=k, act=act, prelu=prelu,
                                        chan_map=chan_map, f32=f32, first=first)))

    def pool(self, src, dst):
        self.steps.append(("pool", src, dst))


def _backbone(p, tail, prelu_names=()):
    """VGG-19 prefix + the net-specific tail; returns the name of the 128-channel feature layer's spec."""
    p.buf("x32", 32, 0)
    p.steps.append(("im2col", "x32"))
    level = 0
    cur = ("x32", 0, 32)
    toggle = 0
    layers = [l for l in _VGG] + list(tail)
    for item in layers:
        if item == "pool":
            level += 1
            name = p.buf("pool%d" % level, cur[2], level)
            p.pool(cur[0], name)
            cur = (name, 0, cur[2])
            continue
        lname, cin, cout = item
        out = p.buf("L%d_%d_%s" % (level, cout, "ab"[toggle]), cout, level)
        if out == cur[0]:
            toggle ^= 1
            out = p.buf("L%d_%d_%s" % (level, cout, "ab"[toggle]), cout, level)
        toggle ^= 1
        if lname in prelu_names:
            p.conv(lname, cur, (out, 0), cout=cout, k=3, act=PRELU, prelu="prelu" + lname[4:], first=(lname == "conv1_1"))
        else:
            p.conv(lname, cur, (out, 0), cout=cout, k=3, act=RELU, first=(lname == "conv1_1"))
        cur = (out, 0, cout)
    return cur


def _emit_feature(p, feat_spec, src, targets, prelu=None):
    """The last backbone layer (-> 128 channels) is written once per buffer that needs `feat` as a slice."""
    lname, cin, cout = feat_spec
    for t in targets:
        p.conv(lname, src, (t, 0), cout=cout, k=3, act=PRELU if prelu else RELU, prelu=prelu)


def build_program(kind):
    p = _Program(kind)
    if kind == "coco":
        src = _backbone(p, [("conv4_3_CPM", 512, 256)])
        # stage input layout [feat 128 | L1 38->40 | L2 19->24]; reference order is cat[L1, L2, feat] (model.py:308)
        for nm in ("catA", "catB"):
            p.buf(nm, 192, 3)
        _emit_feature(p, ("conv4_4_CPM", 256, 128), src, ["catA", "catB"])
        cmap = [57 + i for i in range(128)] + list(range(38)) + [None, None] + [38 + i for i in range(19)] + [None] * 5
        p.buf("t1", 128, 3), p.buf("t2", 128, 3), p.buf("u512", 512, 3)
        p.outputs = [("paf", 38), ("heat", 19)]
        slices = {1: (128, 38), 2: (168, 19)}
        for br in (1, 2):
            off, cout = slices[br]
            p.conv("conv5_1_CPM_L%d" % br, ("catA", 0, 128), ("t1", 0))
            p.conv("conv5_2_CPM_L%d" % br, ("t1", 0, 128), ("t2", 0))
            p.conv("conv5_3_CPM_L%d" % br, ("t2", 0, 128), ("t1", 0))
            p.conv("conv5_4_CPM_L%d" % br, ("t1", 0, 128), ("u512", 0), cout=512, k=1)
            p.conv("conv5_5_CPM_L%d" % br, ("u512", 0, 512), ("catA", off), cout=cout, k=1, act=NONE)
        cur, nxt = "catA", "catB"
        for st in range(2, 7):
            for br in (1, 2):
                off, cout = slices[br]
                tag = "stage%d_L%d" % (st, br)
                p.conv("Mconv1_" + tag, (cur, 0, 192), ("t1", 0), k=7, chan_map=cmap)
                p.conv("Mconv2_" + tag, ("t1", 0, 128), ("t2", 0), k=7)
                p.conv("Mconv3_" + tag, ("t2", 0, 128), ("t1", 0), k=7)
                p.conv("Mconv4_" + tag, ("t1", 0, 128), ("t2", 0), k=7)
                p.conv("Mconv5_" + tag, ("t2", 0, 128), ("t1", 0), k=7)
                p.conv("Mconv6_" + tag, ("t1", 0, 128), ("t2", 0), k=1)
                # model.py:215-218 omits 'Mconv7_stage6_L2' from no_relu_layers: the final heat map keeps its ReLU
                last = st == 6
                act = RELU if (last and br == 2) else NONE
                p.conv("Mconv7_" + tag, ("t2", 0, 128), None if last else (nxt, off), cout=cout, k=1, act=act,
                       f32=(br - 1) if last else None)
            cur, nxt = nxt, cur
    elif kind == "hand":
        src = _backbone(p, [("conv4_3", 512, 512), ("conv4_4", 512, 512), ("conv5_1", 512, 512), ("conv5_2", 512, 512)])
        for nm in ("catA", "catB"):
            p.buf(nm, 152, 3)  # [feat 128 | heat 22->24]; reference order cat[heat, feat] (model.py:397)
        _emit_feature(p, ("conv5_3_CPM", 512, 128), src, ["catA", "catB"])
        cmap = [22 + i for i in range(128)] + list(range(22)) + [None, None]
        p.buf("t1", 128, 3), p.buf("t2", 128, 3), p.buf("u512", 512, 3)
        p.outputs = [("heat", 22)]
        p.conv("conv6_1_CPM", ("catA", 0, 128), ("u512", 0), cout=512, k=1)
        p.conv("conv6_2_CPM", ("u512", 0, 512), ("catA", 128), cout=22, k=1, act=NONE)
        cur, nxt = "catA", "catB"
        for st in range(2, 7):
            p.conv("Mconv1_stage%d" % st, (cur, 0, 152), ("t1", 0), k=7, chan_map=cmap)
            p.conv("Mconv2_stage%d" % st, ("t1", 0, 128), ("t2", 0), k=7)
            p.conv("Mconv3_stage%d" % st, ("t2", 0, 128), ("t1", 0), k=7)
            p.conv("Mconv4_stage%d" % st, ("t1", 0, 128), ("t2", 0), k=7)
            p.conv("Mconv5_stage%d" % st, ("t2", 0, 128), ("t1", 0), k=7)
            p.conv("Mconv6_stage%d" % st, ("t1", 0, 128), ("t2", 0), k=1)
            last = st == 6
            p.conv("Mconv7_stage%d" % st, ("t2", 0, 128), None if last else (nxt, 128), cout=22, k=1, act=NONE,
                   f32=0 if last else None)
            cur, nxt = nxt, cur
    elif kind == "body25":
        src = _backbone(p, [("conv4_3_CPM", 512, 256)], prelu_names=("conv4_2", "conv4_3_CPM"))
        # PAF-stage input [feat 128 | PAF 52->56] (reference cat[feat, PAF], model.py:190);
        # last heat-stage input [feat 128 | heat 26->32 | PAF 52->56] (reference cat[feat, heat, PAF], model.py:199)
        p.buf("catA", 184, 3), p.buf("catB", 184, 3), p.buf("catD", 216, 3)
        _emit_feature(p, ("conv4_4_CPM", 256, 128), src, ["catA", "catB", "catD"], prelu="prelu4_4_CPM")
        map184 = list(range(180)) + [None] * 4
        map216 = list(range(128)) + [128 + i for i in range(26)] + [None] * 6 + [154 + i for i in range(52)] + [None] * 4
        for w in (96, 128):
            p.buf("dense%d_a" % w, 3 * w, 3), p.buf("dense%d_b" % w, 3 * w, 3)
        p.buf("m256", 256, 3), p.buf("m512", 512, 3)
        p.outputs = [("paf", 52), ("heat", 26)]

        def stage(br, st, src_spec, cmap, width, mid, dsts, f32):
            nout = 52 if br == 2 else 26
            tag = "stage%d_L%d" % (st, br)
            cur = src_spec
            bufs = ["dense%d_a" % width, "dense%d_b" % width]
            for blk in range(1, 6):
                y = bufs[(blk - 1) % 2]
                for j in range(3):
                    name = "Mconv%d_%s_%d" % (blk, tag, j)
                    pre = "Mprelu%d_%s_%d" % (blk, tag, j)
                    s = cur if j == 0 else (y, (j - 1) * width, width)
                    p.conv(name, s, (y, j * width), cout=width, k=3, act=PRELU, prelu=pre,
                           chan_map=cmap if (blk == 1 and j == 0) else None)
                cur = (y, 0, 3 * width)
            mbuf = "m%d" % mid
            p.conv("Mconv6_" + tag, cur, (mbuf, 0), cout=mid, k=1, act=PRELU, prelu="Mprelu6_" + tag)
            if not dsts:
                p.conv("Mconv7_" + tag, (mbuf, 0, mid), None, cout=nout, k=1, act=NONE, f32=f32)
            for i, d in enumerate(dsts):
                p.conv("Mconv7_" + tag, (mbuf, 0, mid), d, cout=nout, k=1, act=NONE, f32=f32 if i == 0 else None)

        stage(2, 0, ("catA", 0, 128), None, 96, 256, [("catA", 128)], None)
        stage(2, 1, ("catA", 0, 184), map184, 128, 512, [("catB", 128)], None)
        stage(2, 2, ("catB", 0, 184), map184, 128, 512, [("catA", 128)], None)
        stage(2, 3, ("catA", 0, 184), map184, 128, 512, [("catB", 128), ("catD", 160)], 0)
        stage(1, 0, ("catB", 0, 184), map184, 96, 256, [("catD", 128)], None)
        stage(1, 1, ("catD", 0, 216), map216, 128, 512, [], 1)
    else:
        raise ValueError("unknown network kind %r" % (kind,))
    return p


def _up64(c):
    return (c + 63) // 64 * 64


def pair_followers(steps, si):
    """1x1 conv step si whose output only feeds the 1x1 layer right behind it (Mconv6 -> Mconv7, conv5_4 -> conv5_5,
    conv6_1 -> conv6_2; body25 writes one Mconv7 into two buffers = two steps of the same layer): the step indices of that
    layer, which then runs in the same launch (csrc/conv_umma.cu, variant 6), else []."""
    s = steps[si][1]
    if steps[si][0] != "conv" or s["first"] or s["k"] != 1 or s["dst"] is None:
        return []
    if s["f32"] is not None or s["cout"] % 64 != 0 or not 64 <= s["cout"] <= 512 or s["dst"][1] != 0:
        return []
    want = (s["dst"][0], 0, s["cout"])
    out = []
    sj = si + 1
    while sj < len(steps) and steps[sj][0] == "conv":
        t = steps[sj][1]
        if not (t["k"] == 1 and tuple(t["src"]) == want and t["cout"] <= 64 and (not out or t["layer"] == steps[out[0]][1]["layer"])):
            break
        out.append(sj)
        sj += 1
    if not out or len(out) > 2 or sum(steps[j][1]["f32"] is not None for j in out) > 1:
        return []
    if sum(steps[j][1]["dst"] is not None for j in out) == 0 and all(steps[j][1]["f32"] is None for j in out):
        return []
    # nobody else may read the intermediate before it is written again
    for sk in range(sj, len(steps)):
        if steps[sk][0] != "conv":
            continue
        u = steps[sk][1]
        if not u["first"] and u["src"][0] == want[0]:
            return []
        if u["dst"] is not None and u["dst"][0] == want[0]:
            break
    return out


def find_pairs(steps, enabled=True):
    """{index of the first layer of a chained 1x1 pair: [indices of the second layer's steps]}."""
    if not enabled:
        return {}
    pairs = {si: pair_followers(steps, si) for si in range(len(steps)) if steps[si][0] == "conv"}
    return {si: f for si, f in pairs.items() if f}


class _Instance:
    """One (batch, h, w) instantiation: device buffers + the C launch plan.

    `share` = an instance of the same (h, w) with a batch size >= n: this one then records its plan over the leading n
    images of that instance's buffers (NHWC / NCHW batch prefixes are contiguous) and owns no memory of its own - an
    exact-size replay for the cost of a few hundred tensor-map encodings."""

    def __init__(self, net, n, h, w, share=None, sm_budget=0):
        L = _lib.lib()
        dev = net.device
        self.net = net
        self.n, self.h, self.w = n, h, w
        self.flops_algorithmic = algorithmic_flops(net.kind, n, h, w)
        steps = net.program.steps

        def pool_fusable(si):
            """conv step si followed by a max-pool of its own output (all three pools of all three networks are)."""
            s, nxt = steps[si][1], steps[si + 1] if si + 1 < len(steps) else None
            return (nxt is not None and nxt[0] == "pool" and s["dst"] is not None and nxt[1] == s["dst"][0]
                    and s["k"] >= 3 and s["src"][2] >= 64 and s["cout"] >= 48 and s["f32"] is None and not s["first"])

        pairs = find_pairs(steps, net.tuning.get("pair", True))
        paired = {j for f in pairs.values() for j in f}
        packed_index, ci_ = {}, 0
        for si, step in enumerate(steps):
            if step[0] == "conv":
                packed_index[si] = ci_
                ci_ += 1

        if share is not None:
            if share.h != h or share.w != w or share.n < n or share.net is not net:
                raise ValueError("cannot share buffers of a %dx%dx%d instance for %dx%dx%d" % (share.n, share.h, share.w, n, h, w))
            self.input = share.input[:n]
            self.bufs = {k: v[:n] for k, v in share.bufs.items()}
            self.outputs = [o[:n] for o in share.outputs]
            self._keep = share
        else:
            self.input = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
            self.bufs = {}
            # buffers some launch actually touches (the fused first layer needs no patch buffer, a fused pool no
            # full-size output)
            used = set()
            for si, step in enumerate(steps):
                if step[0] == "conv":
                    s = step[1]
                    if not s["first"] and si not in paired:
                        used.add(s["src"][0])
                    if pool_fusable(si):
                        used.add(steps[si + 1][2])
                    elif s["dst"] is not None and si not in pairs:   # a pair's intermediate stays in shared memory
                        used.add(s["dst"][0])
            for name, (ch, level) in net.program.bufs.items():
                if name not in used:
                    continue
                # physical width rounded up to 64 channels (zeros, never written): every 64-channel box is in bounds
                self.bufs[name] = torch.zeros((n, h >> level, w >> level, _up64(ch) if ch > 32 else ch), dtype=torch.bfloat16,
                                              device=dev)
            gh, gw = h // 8, w // 8
            self.outputs = [torch.empty((n, ch, gh, gw), dtype=torch.float32, device=dev) for _, ch in net.program.outputs]
        handle = C.c_void_p()
        _lib.check(L.islpose_plan_create(C.byref(handle)), "islpose_plan_create")
        self.handle = handle
        fused_pools = set()
        self.op_names = []   # one name per recorded launch (tools/layer_times.py)
        for si, step in enumerate(steps):
            if step[0] == "im2col":
                continue   # conv1_1 gathers its patches itself (csrc/conv_first.cu)
            if step[0] == "pool":
                if si not in fused_pools:
                    raise _lib.IslposeError("max-pool after %r cannot be fused into its producer" % (steps[si - 1],))
                continue   # done in the epilogue of the layer before it
            s = step[1]
            if si in paired:
                continue   # runs inside the launch of the 1x1 layer before it
            wt, bias, slope = net.packed[packed_index[si]]
            if s["first"]:
                # conv1_1 in one launch straight from the float32 network input (csrc/conv_first.cu)
                db = self.bufs[s["dst"][0]]
                _lib.check(L.islpose_plan_add_first_conv(handle, _lib.ptr(self.input), _lib.ptr(wt), _lib.ptr(bias),
                                                         _lib.ptr(slope), C.c_void_p(db.data_ptr() + 2 * s["dst"][1]),
                                                         db.shape[3], n, h, w, 1 if s["act"] == RELU else 0),
                           "islpose_plan_add_first_conv")
                self.op_names.append(s["layer"])
                continue
            sb = self.bufs[s["src"][0]]
            d = _lib.ConvDesc()
            d.in_ = sb.data_ptr() + 2 * s["src"][1]
            d.in_c = s["src"][2]
            d.in_cstride = sb.shape[3]
            d.in_c_readable = sb.shape[3] - s["src"][1]
            d.w_cin = wt.shape[2]
            d.n, d.h, d.w = n, sb.shape[1], sb.shape[2]
            d.weights = wt.data_ptr()
            d.cout = wt.shape[1]
            d.ksize = s["k"]
            d.bias = bias.data_ptr()
            d.slope = slope.data_ptr()
            if pool_fusable(si):
                # nn.MaxPool2d(2, 2) fused into this layer's epilogue: the full-resolution tensor is never written
                db = self.bufs[steps[si + 1][2]]
                d.out_bf16 = db.data_ptr()
                d.out_cstride = db.shape[3]
                d.pool = 1
                fused_pools.add(si + 1)
            elif si in pairs:
                # this 1x1 layer and the 1x1 layer behind it in one launch; the outputs are the second layer's
                fol = [steps[j][1] for j in pairs[si]]
                wt2, bias2, slope2 = net.packed[packed_index[pairs[si][0]]]
                d.weights2, d.cout2 = wt2.data_ptr(), wt2.shape[1]
                d.bias2, d.slope2 = bias2.data_ptr(), slope2.data_ptr()
                dsts = [t["dst"] for t in fol if t["dst"] is not None]
                if dsts:
                    db = self.bufs[dsts[0][0]]
                    d.out2_bf16, d.out2_cstride = db.data_ptr() + 2 * dsts[0][1], db.shape[3]
                if len(dsts) > 1:
                    db = self.bufs[dsts[1][0]]
                    d.out2b_bf16, d.out2b_cstride = db.data_ptr() + 2 * dsts[1][1], db.shape[3]
                for t in fol:
                    if t["f32"] is not None:
                        o = self.outputs[t["f32"]]
                        d.out2_f32, d.out2_f32_channels = o.data_ptr(), o.shape[1]
            elif s["dst"] is not None:
                db = self.bufs[s["dst"][0]]
                d.out_bf16 = db.data_ptr() + 2 * s["dst"][1]
                d.out_cstride = db.shape[3]
            if s["f32"] is not None:
                o = self.outputs[s["f32"]]
                d.out_f32 = o.data_ptr()
                d.out_f32_channels = o.shape[1]
            cfg = net.tuning
            d.n_tile, d.stages = cfg.get("n_tile", 0), cfg.get("stages", 0)
            d.sm_budget = sm_budget
            _lib.check(L.islpose_plan_add_conv(handle, C.byref(d)), "islpose_plan_add_conv(%s)" % s["layer"])
            self.op_names.append(s["layer"] + ("+pool" if d.pool else "") + ("+" + steps[pairs[si][0]][1]["layer"] if si in pairs else ""))
        self.flops = L.islpose_plan_conv_flops(handle)
        self.launches = L.islpose_plan_num_launches(handle)
        if not net.tuning.get("graph", True):
            _lib.check(L.islpose_plan_set_graph(handle, 0), "islpose_plan_set_graph")
        # the zero-fills above ran on the current stream, but the plan may be replayed on any stream: make the
        # buffers (in particular their never-written zero pad channels) globally visible before first use
        if share is None:
            torch.cuda.current_stream().synchronize()

    def run(self):
        _lib.check(_lib.lib().islpose_plan_run(self.handle, _lib.stream_ptr()), "islpose_plan_run")

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().islpose_plan_destroy(self.handle)
        except Exception:
            pass


def algorithmic_flops(kind, n, h, w):
    """sum over layers of 2*Cin*Cout*k*k*Hout*Wout on the unpadded channel counts."""
    p = build_program(kind)
    seen = set()
    total = 0
    for step in p.steps:
        if step[0] != "conv":
            continue
        s = step[1]
        if s["layer"] in seen:
            continue  # a layer written to several buffers is one layer of the network
        seen.add(s["layer"])
        level = p.bufs[s["src"][0]][1]
        cin = 3 if s["first"] else (sum(1 for c in s["chan_map"] if c is not None) if s["chan_map"] else s["src"][2])
        total += 2 * cin * s["cout"] * s["k"] ** 2 * (h >> level) * (w >> level) * n
    return total


class _LayerParams(torch.nn.Module):
    """Holds one Caffe layer's tensors so that the module's state-dict keys are the flat Caffe names of the weight
    files ('conv1_1.weight', 'conv1_1.bias', 'Mprelu1_stage0_L2_0.weight', ...)."""

    def __init__(self, weight, bias=None):
        super().__init__()
        self.weight = torch.nn.Parameter(weight, requires_grad=False)
        if bias is not None:
            self.bias = torch.nn.Parameter(bias, requires_grad=False)


class PoseNet(torch.nn.Module):
    """The reference nn.Module's interface (model.py: bodypose_model / bodypose_25_model / handpose_model) backed by
    launch plans: `model(data float32 [N,3,h,w]) -> (PAF, heat)` / `-> heat`, float32 NCHW on the module's device.

    It is a real torch.nn.Module - float32 parameters under the flat Caffe names, state_dict / load_state_dict /
    named_parameters / eval / to(device) / cuda(device) - so wrappers that only know nn.Module keep working
    (keras.layers.TorchModuleWrapper in ISL_Model_parameter.py:44-47, `.to(device)` in extract_features_mp.py:150,
    `for param in self.parameters()` in model.py:167-168). The kernels read bf16 copies of the parameters packed
    tap-major in buffer channel order; load_state_dict() and to() refresh them. There is no CPU path: to('cpu'),
    half() or double() raise."""

    def __init__(self, kind, flat_weights, device=None, tuning=None):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.IslposeError("PoseNet needs a CUDA device (sm_100a); there is no CPU path")
        _lib.lib()
        self.kind = kind
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        if self.device.type != "cuda":
            raise _lib.IslposeError("PoseNet runs on CUDA devices only, got %s" % (self.device,))
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.program = build_program(kind)
        self.tuning = dict(tuning or {})
        seen = set()
        for step in self.program.steps:
            if step[0] != "conv" or step[1]["layer"] in seen:
                continue
            s = step[1]
            seen.add(s["layer"])
            w = torch.as_tensor(flat_weights[s["layer"] + ".weight"]).detach().to(torch.float32)
            b = torch.as_tensor(flat_weights[s["layer"] + ".bias"]).detach().to(torch.float32)
            cin = 3 if s["first"] else (sum(1 for c in s["chan_map"] if c is not None) if s["chan_map"] else s["src"][2])
            if tuple(w.shape) != (s["cout"], cin, s["k"], s["k"]) or tuple(b.shape) != (s["cout"],):
                raise ValueError("%s: weight shape %s does not match the %s network" % (s["layer"], tuple(w.shape), kind))
            self.add_module(s["layer"], _LayerParams(w.to(self.device), b.to(self.device)))
            if s["prelu"] is not None:
                a = torch.as_tensor(flat_weights[s["prelu"] + ".weight"]).detach().to(torch.float32)
                self.add_module(s["prelu"], _LayerParams(a.to(self.device)))
        self.packed = []
        self._instances = {}
        self.timing = None   # set to a list: Body / Hand append (start, end, flops, launches) per network phase (fork to join)
        self._chan_maps = {}
        self._pack()
        self.eval()

    def _pack(self):
        """float32 parameters -> the kernels' operands: bf16 [tap][cout][cin of the buffer slice, padded to 64], float32
        bias and negative-side slope (0 = ReLU, 1 = none, PReLU weight otherwise), packed on the device
        (csrc/pack.cu). Existing plans keep pointing at the old tensors, so they are dropped."""
        L = _lib.lib()
        self._instances = {}
        packed = []
        with torch.cuda.device(self.device):
            st = _lib.stream_ptr()
            for step in self.program.steps:
                if step[0] != "conv":
                    continue
                s = step[1]
                lp = getattr(self, s["layer"])
                w, b = lp.weight.data, lp.bias.data
                cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
                in_c = s["src"][2]
                w_cin = in_c if (s["first"] or in_c % 64 == 0) else _up64(in_c)
                cmap = None
                if s["chan_map"] is not None:
                    key = tuple(-1 if c is None else c for c in s["chan_map"])
                    if sorted(c for c in key if c >= 0) != list(range(cin)):
                        raise ValueError("channel map of %s does not cover Cin=%d exactly once" % (s["layer"], cin))
                    cmap = self._chan_maps.get(key)
                    if cmap is None:
                        cmap = self._chan_maps[key] = torch.tensor(key, dtype=torch.int32, device=self.device)
                elif not s["first"] and cin != in_c:
                    raise ValueError("%s: slice width %d does not match Cin %d" % (s["layer"], in_c, cin))
                taps = 1 if s["first"] else k * k
                wt = torch.empty((taps, cout, w_cin), dtype=torch.bfloat16, device=self.device)
                _lib.check(L.islpose_pack_conv_weights(_lib.ptr(w.contiguous()), cout, cin, k, _lib.ptr(cmap), in_c, w_cin,
                                                       1 if s["first"] else 0, _lib.ptr(wt), st), "islpose_pack_conv_weights")
                bias = torch.zeros(512, dtype=torch.float32, device=self.device)
                bias[:cout] = b
                slope = torch.zeros(512, dtype=torch.float32, device=self.device)
                if s["act"] == NONE:
                    slope[:cout] = 1.0
                elif s["act"] == PRELU:
                    slope[:cout] = getattr(self, s["prelu"]).weight.data
                packed.append((wt, bias, slope))
            torch.cuda.current_stream().synchronize()
        self.packed = packed

    # ---- nn.Module surface -----------------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        out = super().load_state_dict(state_dict, strict=strict, **kw)
        self._pack()
        return out

    def _apply(self, fn, *args, **kw):
        """to() / cuda() / float() arrive here. The parameters must stay float32 on a CUDA device; moving to another
        GPU re-packs the operands there."""
        probe = fn(torch.empty(0, dtype=torch.float32, device=self.device))
        if probe.device.type != "cuda":
            raise _lib.IslposeError("PoseNet has no CPU path: cannot move it to %s" % (probe.device,))
        if probe.dtype != torch.float32:
            raise _lib.IslposeError("PoseNet keeps float32 parameters (bf16 operands are derived from them), not %s" % (probe.dtype,))
        out = super()._apply(fn, *args, **kw)
        if probe.device != self.device:
            self.device = probe.device
            self._chan_maps = {}
            self._pack()
        return out

    def instance(self, n, h, w, lane=0, exact_of=None, sm_budget=0):
        """The plan (and its activation buffers) for one input shape. `lane` selects an independent copy, so that
        two batches of the same shape can be in flight at once. exact_of = a batch capacity >= n: the plan for n
        images runs inside the buffers of the (capacity, h, w, lane) instance instead of allocating its own.
        sm_budget > 0: the plan's persistent kernels keep to that many SMs (share_sms)."""
        if h % 8 or w % 8:
            raise ValueError("network input must be a multiple of 8 in both dimensions, got %dx%d" % (h, w))
        cap = exact_of if exact_of is not None else n
        if cap != n or sm_budget:
            key = (n, h, w, lane, cap, sm_budget)
            inst = self._instances.get(key)
            if inst is None:
                parent = self.instance(cap, h, w, lane)
                with torch.cuda.device(self.device):
                    inst = _Instance(self, n, h, w, share=parent, sm_budget=sm_budget)
                self._instances[key] = inst
            return inst
        key = (n, h, w, lane)
        inst = self._instances.get(key)
        if inst is None:
            with torch.cuda.device(self.device):
                inst = _Instance(self, n, h, w)
            self._instances[key] = inst
        return inst

    def share_sms(self, shapes):
        """SM budgets for plans that will run side by side, one per (n, h, w) in `shapes` - or all zeros (no limit) when
        at least one of them fills the device on its own. A layer of a small batch has fewer tiles than the device has
        SMs, so a plan tuned for itself spreads every layer thinly over all of them (short, narrow MMAs); four such
        plans on four streams then queue behind each other. With a share each - proportional to its work, at least 4 -
        they run truly concurrently on fuller tiles (measured on a single C2 frame: 5.0 ms -> see DESIGN.md)."""
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        if len(shapes) < 2:
            return [0] * len(shapes)
        tiles = [n * -(-(h // 8) * (w // 8) // 256) for (n, h, w) in shapes]   # 256-pixel tiles on the stride-8 grid
        if max(tiles) >= sms:
            return [0] * len(shapes)
        work = [n * h * w for (n, h, w) in shapes]
        left = sms - 4 * len(shapes)
        budgets = [4 + int(left * wk / sum(work)) for wk in work]
        budgets[work.index(max(work))] += sms - sum(budgets)
        return budgets

    def forward_into(self, data):
        """Runs the plan for `data`'s shape; returns the plan's own float32 output tensors (overwritten by the
        next call with the same shape)."""
        n, c, h, w = data.shape
        with torch.cuda.device(self.device):
            inst = self.instance(n, h, w)
            inst.input.copy_(data)
            inst.run()
        return inst.outputs

    def forward(self, data):
        with torch.no_grad():
            outs = [o.clone() for o in self.forward_into(data.to(self.device, torch.float32))]
        return outs[0] if self.kind == "hand" else (outs[0], outs[1])
