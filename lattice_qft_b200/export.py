"""Mirror of src/export.rs: CSV files byte-compatible with the reference's `csv` + `serde` output.

Host I/O only (no device work).  Formats (computation.rs:129-199, export.rs:7-116):
  results.csv                      header + one ComputationSummary row per run
  energy_data/energy_{i}.csv       headerless, one f64 per row          (read_write_csv(path, false))
  action_data/action_{i}.csv       headerless, one f64 per row
  correlation_data/correlation_{i}.csv   headerless, one row of T f64 per measurement (overwrite_csv)
  plot_data/difference_{i}_{dir}.csv     header x,y,t,value
"""
from __future__ import annotations

import csv
import math
import os
from typing import Iterable, List, Optional, Sequence


def fmt_f64(v: float) -> str:
    """f64 as the csv crate writes it (ryu): shortest round-trip digits, always a '.0' or an
    exponent, exponent form `1e21` / `1e-7` outside [1e-5, 1e16)."""
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "inf" if v > 0 else "-inf"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    r = repr(float(v))
    a = abs(v)
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        e = int(exp)
        if 1e-5 <= a < 1e16:  # ryu stays positional here
            digits = mant.replace("-", "").replace(".", "")
            point = (len(mant.replace("-", "").split(".")[0])) + e
            if point <= 0:
                s = "0." + "0" * (-point) + digits.rstrip("0")
            else:
                s = digits[:point].ljust(point, "0") + "." + (digits[point:].rstrip("0") or "0")
            return ("-" if v < 0 else "") + s
        if mant.endswith(".0"):
            mant = mant[:-2]
        return f"{mant}e{e}"
    if a >= 1e16 or a < 1e-5:
        digits = r.replace("-", "").replace(".", "").lstrip("0")
        e = int(math.floor(math.log10(a)))
        d = digits.rstrip("0") or "0"
        mant = d[0] + ("." + d[1:] if len(d) > 1 else "")
        return ("-" if v < 0 else "") + f"{mant}e{e}"
    return r


def _cell(v) -> str:
    if v is None:
        return ""
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, float):
        return fmt_f64(v)
    return str(v)


def _parse_rows(path: str, has_headers: bool) -> List[List[str]]:
    if not os.path.exists(path):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        open(path, "w").close()  # export.rs:59-63 creates the file
        return []
    with open(path, newline="") as fh:
        rows = [r for r in csv.reader(fh) if r]
    return rows[1:] if has_headers and rows else rows


def _write_rows(path: str, header: Optional[Sequence[str]], rows: Iterable[Sequence[str]]):
    rows = list(rows)
    with open(path, "w", newline="") as fh:
        w = csv.writer(fh, lineterminator="\n")
        if header is not None and rows:  # serde writes the header with the first record
            w.writerow(header)
        w.writerows(rows)


def clean_csv(path: str):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    open(path, "w").close()


def append_floats(path: str, values: Sequence[float]):
    """`Vec<f64>::read_write_csv(path, false)` (export.rs:96-109): append one value per row."""
    rows = _parse_rows(path, False)
    rows += [[fmt_f64(float(v))] for v in values]
    _write_rows(path, None, rows)


def overwrite_float_rows(path: str, table: Sequence[Sequence[float]]):
    """`Vec<Vec<f64>>::overwrite_csv(path)` (export.rs:22-25): one row per measurement."""
    clean_csv(path)
    _write_rows(path, None, [[fmt_f64(float(v)) for v in row] for row in table])


PLOT_HEADER = ["x", "y", "t", "value"]


def append_plotdata(path: str, rows3d: Sequence[Sequence]):
    """`Vec<Plotdata3d>::read_write_csv(path, true)` (export.rs:79-109)."""
    rows = _parse_rows(path, True)
    rows += [[str(x), str(y), str(t), fmt_f64(float(v))] for (x, y, t, v) in rows3d]
    _write_rows(path, PLOT_HEADER, rows)


SUMMARY_HEADER = ["index", "d", "size", "x", "y", "t", "temp", "comptype", "burnin", "iterations", "duration",
                  "algorithm_statistic", "action_data", "energy_data", "difference_data", "correlation_data"]


def read_summaries(path: str) -> List[dict]:
    rows = _parse_rows(path, True)
    return [dict(zip(SUMMARY_HEADER, r)) for r in rows]


def append_summary(path: str, summary: dict):
    rows = _parse_rows(path, True)
    rows.append([_cell(summary.get(k)) for k in SUMMARY_HEADER])
    _write_rows(path, SUMMARY_HEADER, rows)
