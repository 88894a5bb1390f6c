"""CPU: host-side token algebra and partition planning (pure torch, no kernels)."""
import torch

from oracle import imp_oracle as O
from util_hotpath import block_tensors, make_params, rel


def test_fold_query_and_block_tail_equal_reference_attention():
    from imp_b200 import ops
    params = make_params(3)
    in_w, in_b, out_w, out_b, ln_w, ln_b = block_tensors(params, 1)
    g = torch.Generator().manual_seed(0)
    h = torch.relu(torch.randn(257, 256, generator=g))
    c = torch.randn(5, 256, generator=g)
    qt = ops.fold_query(c, in_w, in_b)
    assert rel(qt, O.folded_query(c, in_w, in_b)) < 1e-6
    a = torch.softmax(qt @ h.t(), dim=1)
    out = ops.block_tail(c, a @ h, in_w, in_b, out_w, out_b, ln_w, ln_b)
    blk = dict(zip(["in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "norm1.weight", "norm1.bias"],
                   [in_w, in_b, out_w, out_b, ln_w, ln_b]))
    assert rel(out, O.proto_block(h, c, blk)) < 1e-5


def test_normalize_tokens_is_across_tokens():
    from imp_b200 import modularity as M
    c = torch.randn(2, 7, 256)
    ch = M.normalize_tokens(c)
    assert torch.allclose(ch.norm(dim=1), torch.ones(2, 256), atol=1e-5)
    x = torch.relu(torch.randn(50, 256))
    ref = O.cluster_assignment(x, c[0])
    xh = x / x.norm(dim=1, keepdim=True)
    assert rel(torch.relu(xh @ ch[0].t()), ref) < 1e-6


def test_state_dict_keys_match_reference_names():
    from imp_b200 import model
    net = model.IMPHotPath(n_proto=6, seed=0)
    keys = set(net.state_dict().keys())
    want = {"path_net.0.weight", "path_net.0.bias"}
    for k in range(6):
        want |= {"omic_net.%d.0.weight" % k, "omic_net.%d.0.bias" % k}
    for b in range(2):
        p = "proto_g_blocks.%d." % b
        want |= {p + "cross_attn.in_proj_weight", p + "cross_attn.in_proj_bias", p + "cross_attn.out_proj.weight",
                 p + "cross_attn.out_proj.bias", p + "norm1.weight", p + "norm1.bias"}
    assert keys == want, keys ^ want
    assert net.state_dict()["omic_net.4.0.weight"].shape == (256, 1538)
    assert net.p_proto.shape == (1, 6, 256) and "p_proto" not in keys         # plain tensor, like the reference
    assert net.p_proto.abs().max().item() <= 1 / 6 + 1e-6


def test_shard_bounds_and_slide_assignment():
    from imp_b200 import parallel as P
    for n, w in [(120000, 8), (100, 8), (64, 2), (1, 4), (16384, 3)]:
        b = P.shard_bounds(n, w)
        assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(s % P.TILE == 0 for s, _ in b if s < n)
    lens = [9000, 100, 5000, 4900, 50, 8000, 3000]
    a = P.assign_slides(lens, 3)
    assert sorted(i for r in a for i in r) == list(range(len(lens)))
    loads = [sum(lens[i] for i in r) for r in a]
    assert max(loads) - min(loads) <= max(lens)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) needs no GPU: exactly one JSON line on
    stdout with the contract keys, `impl: reference`, a cpu_baseline describing itself and a zero-copy e2e object."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--patches", "512", "--protos", "16"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "wsi_bags_per_s_fwd_bwd" and d["unit"] == "bags/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_packed_wire_format_round_trips_the_reference_layout():
    """wire.collate_packed on samples shaped like DatasetWrapper_UMEML.__getitem__ (data_manager.py:395-403):
    valid rows only, bf16, offsets; identical (after the bf16 rounding the kernels apply anyway) to the padded
    fp32 reference layout it replaces."""
    from imp_b200 import wire
    g = torch.Generator().manual_seed(0)
    lens = [37, 1, 300]
    samples = []
    for i, n in enumerate(lens):
        bag = torch.randn(n, 512, generator=g)
        samples.append({"img": wire.pad_bag(bag, 512), "mol": torch.rand(3354, generator=g), "label": torch.tensor(i % 4),
                        "survival_month": torch.tensor(12.0 + i), "censorship": torch.tensor(i % 2),
                        "patient_id": "TCGA-%02d" % i, "index": i})
        assert wire.bag_rows(samples[-1]["img"]) == n == O.bag_length(samples[-1]["img"])
    batch = wire.collate_packed(samples)
    assert batch["x_packed"].dtype == torch.bfloat16 and batch["x_packed"].shape == (sum(lens), 512)
    assert batch["cu_seqlens"].tolist() == [0, 37, 38, 338] and batch["max_len"] == 300
    assert batch["omic"].shape == (3, 3354) and batch["patient_id"] == ["TCGA-00", "TCGA-01", "TCGA-02"]
    img = wire.unpack_to_reference_layout(batch, 512)
    ref = torch.stack([s["img"] for s in samples])
    valid = ref != wire.SENTINEL
    assert torch.equal(valid, img != wire.SENTINEL)
    assert torch.equal(img[valid], ref[valid].bfloat16().float())
    means = wire.omic_means([{"mol": batch["mol"][:2]}, {"mol": batch["mol"][2:]}])
    assert torch.allclose(means, batch["mol"].mean(0), atol=1e-6)          # trainer.py:286-291
    big = wire.pad_bag(torch.randn(20, 512, generator=g), 16)              # >= target rows: returned unpadded
    assert big.shape == (20, 512) and wire.bag_rows(big) == 20
