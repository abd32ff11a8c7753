"""bench.py pieces that run without a GPU: the FLOP count behind `roofline.achieved`, the host-thread rule of the reference arm and
the JSON line of `--impl reference` (timed here on the 2-block test geometry instead of ViT-H; same code path)."""
import argparse
import importlib
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture()
def bench(monkeypatch):
    saved = {k: v for k, v in sys.modules.items() if k == "segment_anything" or k.startswith("segment_anything.")}
    path = list(sys.path)
    mod = importlib.import_module("bench")
    yield mod
    # the reference arm puts oracle/_ref/GD first on sys.path and imports ITS segment_anything: undo both for the other tests
    for k in [k for k in sys.modules if k == "segment_anything" or k.startswith("segment_anything.")]:
        del sys.modules[k]
    sys.modules.update(saved)
    sys.path[:] = path


def test_encoder_gemm_flops_match_the_survey(bench):
    from samrs_b200.config import geometry
    # 5 169.6 GFLOP of GEMMs per ViT-H encode (patch embed + 32 blocks + neck)
    assert bench.gemm_flops_per_encode(geometry("vit_h")) / 1e9 == pytest.approx(5169.6, abs=0.5)


def test_host_threads_is_capped_by_override_and_affinity(bench, monkeypatch):
    monkeypatch.setenv("SAMRS_REF_THREADS", "2")
    assert bench.host_threads() == min(2, len(os.sched_getaffinity(0)))
    monkeypatch.delenv("SAMRS_REF_THREADS")
    n = bench.host_threads()
    assert 1 <= n <= len(os.sched_getaffinity(0))


def test_reference_arm_prints_the_contract_line(bench, monkeypatch, capsys):
    monkeypatch.setattr(bench, "VARIANT", "vit_t64")
    monkeypatch.setenv("SAMRS_REF_THREADS", "4")
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=1, warmup=1, impl="reference", config="hbox32", no_cpu_baseline=False, no_extra=False)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "box_prompted_masks_per_sec" and line["unit"] == "masks/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["n_gpus"] == 1
    assert line["steps"] >= 3 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["value"] == pytest.approx(32 / (line["ms_per_step"] / 1000.0))
    assert "32 hbox prompts" in line["config"]["workload"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == min(4, len(os.sched_getaffinity(0))) and cb["value"] == line["value"]
    staged = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "GD", "segment_anything"))
    assert cb["kind"] == ("reference" if staged else "port")
    assert line["e2e"] == {"value": line["value"], "unit": "masks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only(bench, monkeypatch, capsys):
    monkeypatch.setenv("RANK", "1")
    args = argparse.Namespace(gpus=2, steps=1, warmup=1, impl="reference", config="hbox32", no_cpu_baseline=False, no_extra=False)
    bench.run_reference(args)
    assert capsys.readouterr().out == ""


def test_roofline_extras_use_the_survey_flop_counts(bench):
    from samrs_b200.config import geometry
    g = geometry("vit_h")
    fw, fg = bench.attention_flops_per_encode(g)
    assert fw / 1e9 == pytest.approx(115.09, abs=0.05) and fg / 1e9 == pytest.approx(343.60, abs=0.05)
    prof = {"attn_window": (1.6, 56), "attn_global": (1.4, 8), "gemm_tc": (11.0, 262)}
    ex = bench.roofline_extras(g, prof, prof_steps=2, masks_per_s_per_gpu=3840.0, prompts=32, tokens_per_prompt=7, peak_tf=1456.2)
    assert ex["attn_window"]["achieved"] == pytest.approx(115.09 * 2 / 1.6, rel=1e-3)             # GFLOP / ms = TFLOP/s
    assert ex["attn_global"]["achieved"] == pytest.approx(343.60 * 2 / 1.4, rel=1e-3)
    assert ex["attn_window"]["launches_per_step"] == 28 and ex["attn_global"]["ms_per_step"] == pytest.approx(0.7)
    # 3 840 masks/s = 120 tiles/s x 5 757.7 GFLOP per tile (179.9 GFLOP per mask)
    assert ex["whole_step"]["gflop_per_tile"] / 32 == pytest.approx(179.9, abs=0.05)
    assert ex["whole_step"]["achieved"] == pytest.approx(3840 * 179.93 / 1e3, rel=1e-3)
    assert ex["whole_step"]["frac"] == pytest.approx(ex["whole_step"]["achieved"] / 1456.2)
    # unknown token count (5-point prompts carry 11 tokens) or another geometry: the whole-step reading is left out, nothing raises
    assert "whole_step" not in bench.roofline_extras(g, prof, 2, 3840.0, 32, 11, 1456.2)
    assert "whole_step" not in bench.roofline_extras(geometry("vit_t64"), {}, 2, 100.0, 8, 7, 1456.2)
