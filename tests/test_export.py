"""CSV formats of src/export.rs / parse_simulation_results (computation.rs:129-199): host I/O, no GPU."""
import os

import pytest

from lattice_qft_b200 import export


@pytest.mark.parametrize("v,s", [
    (682.6452578125001, "682.6452578125001"), (0.2, "0.2"), (1.0, "1.0"), (0.0, "0.0"), (-3.0, "-3.0"),
    (1e21, "1e21"), (1e-7, "1e-7"), (1e16, "1e16"), (1e15, "1000000000000000.0"), (0.00001, "0.00001"),
    (1.5e-6, "1.5e-6"), (-3.25e20, "-3.25e20"), (123456789012345680.0, "1.2345678901234568e17"), (0.21, "0.21"),
])
def test_f64_format_is_ryu(v, s):
    assert export.fmt_f64(v) == s


def test_reference_results_header_matches(tmp_path):
    """Header of the reference's own data/results.csv (first line of the shipped file)."""
    want = ("index,d,size,x,y,t,temp,comptype,burnin,iterations,duration,algorithm_statistic,action_data,"
            "energy_data,difference_data,correlation_data")
    assert ",".join(export.SUMMARY_HEADER) == want
    ref = "/root/reference/data/results.csv"
    if os.path.exists(ref):  # only in the build container
        with open(ref) as fh:
            assert fh.readline().strip() == want
            # a shipped row round-trips through our writer byte for byte
            row = fh.readline().strip()
        path = str(tmp_path / "r.csv")
        fields = row.split(",")
        summary = dict(zip(export.SUMMARY_HEADER, fields))
        summary["temp"] = float(summary["temp"])
        summary["action_data"], summary["energy_data"] = summary["action_data"] == "true", summary["energy_data"] == "true"
        summary["difference_data"] = summary["difference_data"] == "true"
        summary["correlation_data"] = summary["correlation_data"] == "true"
        summary["algorithm_statistic"] = None
        export.append_summary(path, summary)
        assert open(path).read().splitlines() == [want, row]


def test_append_semantics(tmp_path):
    p = str(tmp_path / "energy_data" / "energy_3.csv")
    os.makedirs(os.path.dirname(p))
    export.append_floats(p, [1.0, 2.5])
    export.append_floats(p, [682.6452578125001])
    assert open(p).read() == "1.0\n2.5\n682.6452578125001\n"           # headerless, one value per row
    c = str(tmp_path / "correlation_0.csv")
    export.overwrite_float_rows(c, [[1.0, 2.0], [3.0, 4.5]])
    export.overwrite_float_rows(c, [[9.0, 8.0]])                       # overwrite_csv clears first
    assert open(c).read() == "9.0,8.0\n"
    d = str(tmp_path / "difference_0_1.csv")
    export.append_plotdata(d, [(0, 0, 0, 0.5), (1, 0, 0, -1.0)])
    assert open(d).read() == "x,y,t,value\n0,0,0,0.5\n1,0,0,-1.0\n"
    r = str(tmp_path / "results.csv")
    for i in range(2):
        export.append_summary(r, {"index": i, "d": 3, "size": 8, "x": 2, "y": 2, "t": 2, "temp": 0.25,
                                  "comptype": "Metropolis Simulation", "burnin": 10, "iterations": 20, "duration": 0,
                                  "algorithm_statistic": None, "action_data": False, "energy_data": True,
                                  "difference_data": False, "correlation_data": False})
    lines = open(r).read().splitlines()
    assert len(lines) == 3 and lines[2] == "1,3,8,2,2,2,0.25,Metropolis Simulation,10,20,0,,false,true,false,false"
    assert [row["index"] for row in export.read_summaries(r)] == ["0", "1"]
