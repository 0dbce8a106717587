"""results_csv.write_header / write_result against the text the reference's test3.py writes (tests/golden/qsim_results.csv,
oracle/gen_golden_sim.py)."""
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "qsim_results.csv")


def test_csv_matches_reference_text(tmp_path):
    from polarcub_b200 import ProbResult as PR
    from polarcub_b200 import results_csv as rc
    path = str(tmp_path / "r.csv")
    rc.write_header(path)
    rc.write_header(path)  # a second call checks the header and writes nothing
    rc.write_result(path, 2, 0.05, None, 0.7136, 6, 64, 100, "TalVardy", 32, 0.46875, 4, 0.125, 0.01, 0.4, 123.5, 24,
                    [PR.SuccessActualIsMax] * 20 + [PR.FailActualWithinRange] * 3 + [PR.SuccessActualSmallerThanMax])
    rc.write_result(path, 3, 0.02, None, 1.4, 5, 32, 100, "TalVardy", 16, 0.73, 9, 0.0, 0.0, 0.7, 1.25, 10,
                    [PR.SuccessActualIsMax] * 10)
    assert open(path).read() == open(GOLD).read()
    bad = str(tmp_path / "bad.csv")
    open(bad, "w").write("x,y\n")
    with pytest.raises(AssertionError):
        rc.write_header(bad)
    assert abs(rc.calc_theoretic_key_rate(3, qer=0.02) - 1.4235) < 1e-3
