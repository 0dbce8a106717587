"""The CSV results file of the reference's information-reconciliation sweeps (test3.py:282-312): same header, same row
layout, so files written by either implementation can be appended to and read by the other (plot.py reads them by column
name)."""
import csv
import math
from collections import Counter

from .QaryPolarEncoderDecoder import ProbResult

HEADER = ["q", "qer", "snr", "theoreticKeyRate", "n", "N", "L", "frozenBitsAlgorithm", "numInfoQudits", "rate", "maxListSize",
          "frameErrorProb", "symbolErrorProb", "keyRate", "yield", "efficiency", "timeRate", "numTrials"]


def write_header(file_name):
    """test3.py:282-295: writes the header into a new file; an existing file must start with exactly this header."""
    header = HEADER + [r.name for r in ProbResult]
    try:
        with open(file_name, 'r') as f:
            for row in f:
                assert row.rstrip('\n').split(",") == header
                return
    except FileNotFoundError:
        with open(file_name, 'a', newline='') as f:
            csv.writer(f).writerow(header)
    except AssertionError:
        raise AssertionError(f"Header of {file_name} is bad.")


def write_result(file_name, q, qer, snr, theoretic_key_rate, n, N, L, frozenBitsAlgorithm, numInfoQudits, rate, maxListSize,
                 frame_error_prob, symbol_error_prob, key_rate, time_rate, numTrials, prob_result_list, verbosity=False):
    """test3.py:297-312: one row per configuration; yield / efficiency only for q = 2; the ProbResult columns are the
    fractions of trials per outcome."""
    if verbosity:
        print("writing results")
    with open(file_name, 'a', newline='') as f:
        writer = csv.writer(f)
        if q == 2:
            yld = (1 - frame_error_prob) * numInfoQudits * math.log(q, 2)
            efficiency = numInfoQudits * math.log(q, 2) / (-qer * math.log(qer, 2) - (1 - qer) * math.log(1 - qer, 2))
        else:
            yld = None
            efficiency = None
        counter = Counter(prob_result_list)
        if verbosity:
            print(counter)
        stats = [counter[r] / numTrials for r in ProbResult]
        writer.writerow([q, qer, snr, theoretic_key_rate, n, N, L, frozenBitsAlgorithm, numInfoQudits, rate, maxListSize,
                         frame_error_prob, symbol_error_prob, key_rate, yld, efficiency, time_rate, numTrials] + stats)


def calc_theoretic_key_rate(q, channel_type="QSC", qer=None, snr=None, rate=None):
    """test3.py:314-322."""
    if channel_type != "QSC":
        raise ValueError("TODO in the reference: channel type " + str(channel_type))
    if qer == 0.0:
        return math.log(q, 2)
    if qer == 1.0:
        return math.log(q / (q - 1), 2)
    return math.log(q, 2) + (1 - qer) * math.log(1 - qer, 2) + qer * math.log(qer / (q - 1), 2)
