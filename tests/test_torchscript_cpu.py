"""`-m model.pt`: the library reads TorchScript archives natively (herro_b200/csrc/torchscript.cpp: ZIP + pickle, no libtorch) where
the reference calls tch::CModule::load_on_device (src/inference.rs:185).  Host-only checks through hb_inspect_model."""
import os
import zipfile

import pytest

import helpers
from herro_b200 import api, weights as hbw

torch = pytest.importorskip("torch")
HB_ERR_MODEL = -3


def _net(seed=3, cfg=None):
    from oracle import forward_ref
    cfg = cfg or hbw.NetConfig()
    return cfg, forward_ref.from_weights(cfg, hbw.random_weights(cfg, seed))


def test_scripted_archive_equals_the_blob(tmp_path):
    blob = helpers.model_path(seed=3)
    cfg, net = _net(seed=3)
    pt = str(tmp_path / "model.pt")
    torch.jit.script(net).save(pt)
    d_blob, h_blob = api.inspect_model(blob)
    d_pt, h_pt = api.inspect_model(pt)
    assert d_blob == d_pt == hbw.config_dict(cfg)
    assert h_blob == h_pt   # every canonical tensor bit-identical, names included


def test_other_sizes_and_a_plain_state_dict(tmp_path):
    cfg = hbw.NetConfig(stem_k=17, channels=128, heads=8, layers=3, ffn=256, collapse=128)
    _, net = _net(seed=5, cfg=cfg)
    blob = helpers.model_path(seed=5, cfg=cfg)
    pt = str(tmp_path / "m.pt")
    torch.jit.script(net).save(pt)
    assert api.inspect_model(pt) == api.inspect_model(blob)
    # torch.save of the state_dict (zipfile serialisation): same parameter names; `heads` is not in a state_dict (default 4)
    sd = str(tmp_path / "sd.pt")
    torch.save({"state_dict": net.state_dict()}, sd)
    d, h = api.inspect_model(sd)
    assert {k: v for k, v in d.items() if k != "heads"} == {k: v for k, v in hbw.config_dict(cfg).items() if k != "heads"}
    assert h == api.inspect_model(blob)[1]


def test_batchnorm_after_the_stem_is_folded(tmp_path):
    cfg, net = _net(seed=7)
    C = cfg.channels
    g = torch.Generator().manual_seed(1)
    sd = dict(net.state_dict())
    bn = {"stem_bn.weight": 1 + 0.1 * torch.randn(C, generator=g), "stem_bn.bias": 0.1 * torch.randn(C, generator=g),
          "stem_bn.running_mean": 0.1 * torch.randn(C, generator=g), "stem_bn.running_var": 1 + 0.1 * torch.rand(C, generator=g),
          "stem_bn.num_batches_tracked": torch.tensor(7)}
    p = str(tmp_path / "bn.pt")
    torch.save({**sd, **bn}, p)
    import sys
    sys.path.insert(0, os.path.join(helpers.ROOT, "tools"))
    import export_weights
    dims, T = export_weights.state_dict_to_tensors({**sd, **bn})
    ref = str(tmp_path / "ref.hbw")
    hbw.save_blob(ref, hbw.NetConfig(heads=4, **dims), T)
    # same fp32 operations in the same order as the numpy fold of tools/export_weights.py: bit-identical
    assert api.inspect_model(p) == api.inspect_model(ref)


def test_non_contiguous_and_half_tensors(tmp_path):
    cfg, net = _net(seed=9)
    sd = dict(net.state_dict())
    sd["layers.0.qkv.weight"] = sd["layers.0.qkv.weight"].t().contiguous().t()      # same values, strides (1, 3C)
    sd["lnf.bias"] = sd["lnf.bias"].double()
    p = str(tmp_path / "nc.pt")
    torch.save(sd, p)
    assert not sd["layers.0.qkv.weight"].is_contiguous()
    assert api.inspect_model(p)[1] == api.inspect_model(helpers.model_path(seed=9))[1]


def test_foreign_graphs_and_damaged_archives_are_rejected(tmp_path):
    from herro_b200.api import HerroError
    p = str(tmp_path / "lin.pt")
    torch.jit.script(torch.nn.Linear(4, 4)).save(p)
    with pytest.raises(HerroError) as ei:
        api.inspect_model(p)
    assert ei.value.code == HB_ERR_MODEL and "stem.weight" in str(ei.value)   # HB_ERR_MODEL, names the first missing parameter
    # compressed entries: never written by PyTorch, refused rather than inflated
    cfg, net = _net(seed=3)
    good = str(tmp_path / "good.pt")
    torch.jit.script(net).save(good)
    comp = str(tmp_path / "comp.pt")
    with zipfile.ZipFile(good) as zi, zipfile.ZipFile(comp, "w", zipfile.ZIP_DEFLATED) as zo:
        for it in zi.infolist():
            zo.writestr(it.filename, zi.read(it.filename))
    with pytest.raises(HerroError) as ei:
        api.inspect_model(comp)
    assert ei.value.code == HB_ERR_MODEL and "compressed" in str(ei.value)
    # truncations of a good archive and of its pickle: an error, never a crash
    raw = open(good, "rb").read()
    for cut in (10, 100, len(raw) // 2, len(raw) - 30):
        t = str(tmp_path / f"cut{cut}.pt")
        open(t, "wb").write(raw[:cut])
        with pytest.raises(HerroError):
            api.inspect_model(t)
    # a tensor whose strides reach outside its storage
    with zipfile.ZipFile(good) as zi:
        names = zi.namelist()
        pk = [n for n in names if n.endswith("data.pkl")][0]
        data = zi.read(pk)
        bad = str(tmp_path / "bad.pt")
        with zipfile.ZipFile(bad, "w", zipfile.ZIP_STORED) as zo:
            for n in names:
                b = zi.read(n)
                if n.endswith("/data/0"):
                    b = b[: len(b) // 2]       # storage 0 (read_pos) shorter than its tensor claims
                zo.writestr(n, b)
    with pytest.raises(HerroError) as ei:
        api.inspect_model(bad)
    assert "outside its storage" in str(ei.value)


def test_mutated_archives_never_crash_the_reader(tmp_path):
    """A model file is untrusted input: random byte flips in the pickle, the central directory or anywhere, and truncations, must end
    in a result or HB_ERR_MODEL (the same corpus was run under ASan/UBSan against torchscript.cpp when it was written)."""
    import random
    from herro_b200.api import HerroError
    cfg, net = _net(seed=3)
    good = str(tmp_path / "good.pt")
    torch.jit.script(net).save(good)
    raw = open(good, "rb").read()
    with zipfile.ZipFile(good) as z:
        pk = [i for i in z.infolist() if i.filename.endswith("data.pkl")][0]
        cd = z.start_dir
    rnd = random.Random(11)
    p = str(tmp_path / "m.pt")
    outcomes = {"ok": 0, "rejected": 0}
    for _ in range(250):
        b = bytearray(raw)
        region = rnd.choice(["pkl", "cd", "any", "trunc"])
        if region == "trunc":
            b = b[: rnd.randrange(len(b))]
        else:
            for _ in range(rnd.choice([1, 2, 4, 8])):
                pos = {"pkl": rnd.randrange(pk.header_offset, pk.header_offset + pk.file_size + 200), "cd": rnd.randrange(cd, len(b)),
                       "any": rnd.randrange(len(b))}[region]
                b[pos] = rnd.randrange(256)
        open(p, "wb").write(b)
        try:
            api.inspect_model(p)
            outcomes["ok"] += 1
        except HerroError as e:
            assert e.code == HB_ERR_MODEL
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 50 and outcomes["ok"] > 20
