"""Metrics in the OpenVM-1 ("V1", FRI prover) JSON schema the reference's tooling consumes
(/root/reference/openvm/metrics-viewer/CLAUDE.md:40-78: {"counter": [...], "gauge": [...]}, each entry {labels: [[k, v]...], metric,
value}; consumer: /root/reference/openvm-riscv/scripts/basic_metrics.py:24-90 via metrics_utils.load_metrics_dataframes).
`segment_metrics` maps pb_last_stage_ms onto the V1 gauge names so that `basic_metrics.py summary-table` runs on our output."""
import json


def segment_metrics(stage_ms, rows, main_cols, perm_cols, n_constraints, n_interactions, segment=0, air_name="PowdrAir", air_id=0,
                    trace_gen_ms=0.0, query_ms=0.0):
    """stage_ms: dict from Context.last_stage_ms(); perm_cols: base columns of the LogUp permutation trace"""
    g = lambda k: float(stage_ms.get(k, 0.0))
    excl = g("total") + query_ms
    gauges = {
        "main_trace_commit_time_ms": g("h2d") + g("lde") + g("merkle"),
        "generate_perm_trace_time_ms": g("logup_gen"),
        "perm_trace_commit_time_ms": g("logup_commit"),
        "quotient_poly_compute_time_ms": g("quotient"),
        "quotient_poly_commit_time_ms": g("qlde") + g("qmerkle"),
        "pcs_opening_time_ms": g("open") + g("fri") + g("pow") + query_ms,
        "stark_prove_excluding_trace_time_ms": excl,
        "trace_gen_time_ms": float(trace_gen_ms),
        "total_proof_time_ms": excl + float(trace_gen_ms),
        "execute_preflight_time_ms": 0.0,
        "execute_metered_time_ms": 0.0,
    }
    seg = [["group", "app_proof"], ["segment", str(segment)]]
    air = seg + [["air_name", air_name], ["air_id", str(air_id)]]
    cells = rows * (main_cols + perm_cols)
    counters = [
        (air, "rows", rows), (air, "main_cols", main_cols), (air, "prep_cols", 0), (air, "perm_cols", perm_cols), (air, "cells", cells),
        (seg, "total_cells", cells), (seg, "total_cells_used", cells), (seg, "main_cells_used", rows * main_cols),
        ([["air_name", air_name], ["air_id", str(air_id)]], "constraints", n_constraints),
        ([["air_name", air_name], ["air_id", str(air_id)]], "interactions", n_interactions),
        (seg, "quotient_deg", 2), (seg, "fri.log_blowup", 1),
    ]
    return {"counter": [{"labels": l, "metric": m, "value": str(v)} for l, m, v in counters],
            "gauge": [{"labels": seg, "metric": m, "value": str(v)} for m, v in gauges.items()]}


def merge(metric_dicts):
    out = {"counter": [], "gauge": []}
    for d in metric_dicts:
        out["counter"] += d["counter"]
        out["gauge"] += d["gauge"]
    return out


def write(path, metrics):
    with open(path, "w") as f:
        json.dump(metrics, f)
