"""CPU checks of the network programs (no kernels): topology, channel maps, weight packing, FLOP accounting."""
import numpy as np
import pytest
import torch

import isl_b200  # noqa: F401
from isl_b200 import nets
from oracle import openpose_oracle as O
from packref import pack_reference


@pytest.mark.parametrize("kind,layers,gflop", [("coco", 92, 271.87), ("body25", 114, 161.15), ("hand", 52, 206.38)])
def test_program_covers_every_layer_and_flops_match_survey(kind, layers, gflop):
    p = nets.build_program(kind)
    names = {s[1]["layer"] for s in p.steps if s[0] == "conv"}
    assert names == {l[0] for l in O.net_layers(kind)} and len(names) == layers
    assert abs(nets.algorithmic_flops(kind, 1, 368, 368) / 1e9 - gflop) < 0.01


@pytest.mark.parametrize("kind", ["coco", "body25", "hand"])
def test_weight_packing_is_a_permutation_of_the_reference_weights(kind):
    """The packing rule (tests/packref.py; the device kernel csrc/pack.cu is compared with it bit for bit on the GPU)."""
    flat = O.make_flat_weights(kind, seed=0)
    spec = {l[0]: l for l in O.net_layers(kind)}
    for step in nets.build_program(kind).steps:
        if step[0] != "conv":
            continue
        s = step[1]
        w = flat[s["layer"] + ".weight"]
        _, cin, cout, k, act, prelu = spec[s["layer"]]
        assert s["cout"] == cout and s["k"] == k and s["act"] == act and s["prelu"] == prelu
        packed = pack_reference(w, s["chan_map"], s["src"][2], s["first"]).float()
        assert packed.shape[1] == cout and packed.shape[2] % 8 == 0 and packed.shape[2] >= s["src"][2]
        # every reference weight appears exactly once (bf16-rounded), the rest is zero
        ref = w.to(torch.bfloat16).float()
        assert torch.isclose(packed.abs().sum(), ref.abs().sum(), rtol=1e-5)
        if s["first"]:
            assert packed.shape == (1, 64, 32) and torch.equal(packed[0, :, 27:], torch.zeros(64, 5))
        elif s["chan_map"] is not None:
            cmap = s["chan_map"]
            used = [c for c in cmap if c is not None]
            assert sorted(used) == list(range(cin))
            i = next(i for i, c in enumerate(cmap) if c == cin - 1)
            assert torch.equal(packed[:, :, i], ref.permute(2, 3, 0, 1).reshape(k * k, cout, cin)[:, :, cin - 1])
            pads = [i for i, c in enumerate(cmap) if c is None]
            assert torch.count_nonzero(packed[:, :, pads]) == 0


def test_slices_are_16_byte_aligned_and_inside_their_buffers():
    for kind in ("coco", "body25", "hand"):
        p = nets.build_program(kind)
        for step in p.steps:
            if step[0] != "conv":
                continue
            s = step[1]
            ch, _ = p.bufs[s["src"][0]]
            assert s["src"][1] % 8 == 0 and s["src"][2] % 8 == 0 and s["src"][1] + s["src"][2] <= ch
            if s["dst"] is not None:
                dch, _ = p.bufs[s["dst"][0]]
                assert s["dst"][1] % 8 == 0 and s["dst"][1] + (s["cout"] + 7) // 8 * 8 <= dch


def test_package_weight_generator_matches_the_oracle_generator():
    """bench.py and the tools draw their seeded weights from isl_b200.synth (the product side may not import oracle/);
    the tests draw theirs from the oracle. Same seeds must mean the same tensors, or goldens and benchmarks diverge."""
    import torch

    from isl_b200 import synth
    from oracle import openpose_oracle as O

    for kind in ("coco", "body25", "hand"):
        for init in ("torch", "he"):
            a = synth.make_flat_weights(kind, seed=3, init=init)
            b = O.make_flat_weights(kind, seed=3, init=init)
            assert list(a) == list(b)
            assert all(torch.equal(a[k], b[k]) for k in a)


def test_chained_1x1_pairs_found_by_the_plan_builder():
    """Every 1x1 layer whose only consumer is the 1x1 layer behind it runs in that layer's launch (csrc/conv_umma.cu variant
    6): Mconv6 -> Mconv7 of every stage, conv5_4 -> conv5_5 (coco), conv6_1 -> conv6_2 (hand); body25's stage that writes its
    Mconv7 into two buffers keeps both destinations; nothing else is paired."""
    from isl_b200 import nets
    want = {"coco": 12, "hand": 6, "body25": 6}
    for kind, n_pairs in want.items():
        steps = nets.build_program(kind).steps
        pairs = nets.find_pairs(steps)
        assert len(pairs) == n_pairs
        assert nets.find_pairs(steps, enabled=False) == {}
        for first, followers in pairs.items():
            a = steps[first][1]
            assert a["k"] == 1 and a["cout"] % 64 == 0 and a["f32"] is None
            names = {steps[j][1]["layer"] for j in followers}
            assert len(names) == 1 and 1 <= len(followers) <= 2
            for j in followers:
                b = steps[j][1]
                assert b["k"] == 1 and tuple(b["src"]) == (a["dst"][0], 0, a["cout"]) and b["cout"] <= 64
            assert a["layer"].replace("Mconv6", "Mconv7").replace("conv5_4", "conv5_5").replace("conv6_1", "conv6_2") in names
        two = [f for f in pairs.values() if len(f) == 2]
        assert len(two) == (1 if kind == "body25" else 0)
        # every 1x1 layer that reads a whole temporary is inside a pair
        paired = {j for f in pairs.values() for j in f}
        for si, st in enumerate(steps):
            if st[0] == "conv" and st[1]["k"] == 1 and not st[1]["first"] and st[1]["cout"] <= 64:
                assert si in paired, st[1]["layer"]
