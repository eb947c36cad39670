"""The multi-GPU product path (as_create_multi): the _host entry points and the two programs sharded over several GPUs of
one box give byte-identical results to one GPU.  Needs at least two GPUs (skipped otherwise; the driver's scaling run and
`gpurun --gpus 2` exercise it)."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import aseq_io, synth
from tests import golden_util as gu

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "amplisolve_b200" / "bin"


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def devices():
    n = n_gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return list(range(min(n, 4)))


@pytest.mark.parametrize("tile", [0, 256])
def test_multi_device_host_entry_points_equal_single_device(ctx, devices, tile):
    from amplisolve_b200 import Context, pack_counts, twin_links
    _, slots, pos_id, U = synth.make_panel(90, seed=311, overlap_frac=0.5, amp_len=(30, 90))
    pos_id = pos_id.copy()
    pos_id[-2] = pos_id[7]          # a group that spans the whole panel: it pins every shard boundary behind it
    P = len(slots)
    normals, ref = synth.make_counts(17, P, depth=2500, seed=311, pos_id=pos_id, ragged_twins=True)
    tumours, _ = synth.make_counts(23, P, depth=2500, seed=312, ref=ref, pos_id=pos_id, somatic_rate=0.02)
    nxt, head = twin_links(pos_id)
    with Context(devices) as multi:
        for c in (ctx, multi):
            c.set_host_tile_slots(tile)
        try:
            for fmt in ("u32", "packed"):
                n_in, n_w = (normals, None) if fmt == "u32" else pack_counts(normals)
                t_in, t_w = (tumours, None) if fmt == "u32" else pack_counts(tumours)
                one = ctx.estimate_thresholds(n_in, 0.002, 100, nxt, head, wide_records=n_w, with_view=True)
                many = multi.estimate_thresholds(n_in, 0.002, 100, nxt, head, wide_records=n_w, with_view=True)
                for k in one:
                    assert one[k].tobytes() == many[k].tobytes(), (fmt, k)
                a = ctx.call_variants(t_in, ref, one["thr_view"], 100, wide_records=t_w)
                b = multi.call_variants(t_in, ref, one["thr_view"], 100, wide_records=t_w)
                assert len(a) > 50 and a.tobytes() == b.tobytes(), fmt
                # a capacity that overflows reports the true count on both
                small = multi.call_variants(t_in, ref, one["thr_view"], 100, cap=10, wide_records=t_w)
                assert small.tobytes() == a.tobytes()
        finally:
            ctx.set_host_tile_slots(0)


def test_programs_on_several_gpus_are_byte_identical_to_one_gpu(tmp_path, devices):
    case = gu.load("synth_small")
    outs = {}
    for tag, devs in (("one", "0"), ("many", ",".join(map(str, devices))), ("widened", None)):
        wd = tmp_path / tag
        wd.mkdir()
        slots = aseq_io.stage_case(wd, case)
        aseq_io.write_fasta(wd, slots, list(case["ref_letters"]))
        # "widened": no device list -- the program starts on GPU 0 and moves to every visible GPU once it knows the job is
        # large (the threshold is lowered to 1 record here), parsing and calling in groups of two samples
        env = dict(os.environ, AS_DEVICES=devs) if devs else dict(os.environ, AS_WIDEN_RECORDS="1", AS_GROUP_SAMPLES="2")
        env.pop("AS_DEVICES", None) if devs is None else None
        r = subprocess.run([str(BIN / "AmpliSolveErrorEstimation"), "panel_design=panel.bed", "reference_genome=ref.fa", "germline_dir=N",
                            f"C_value={float(case['c_value']):.4f}", f"coverage_cutoff={int(case['cutoff'])}", "default_error=0.01",
                            "output_dir=o"], cwd=wd, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout[-2000:]
        table = next((wd / "o").glob("positionSpecificNoise_0*.txt"))
        r = subprocess.run([str(BIN / "AmpliSolveVariantCalling"), f"errorFile=o/{table.name}", "tumour_dir=T", "output_dir=v",
                            f"coverage_cutoff={int(case['cutoff'])}", "p_value=0.05"], cwd=wd, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout[-2000:]
        vcfs = sorted((wd / "v").glob("*.vcf"))
        outs[tag] = (table.read_bytes(), (wd / "v" / "Summary_Variant_Info.txt").read_bytes(),
                     [b"".join(l for l in open(v, "rb") if not l.startswith(b"##fileDate=")) for v in vcfs])
    assert outs["one"] == outs["many"] == outs["widened"]
    assert outs["one"][0].decode() == case["noise_table"]
