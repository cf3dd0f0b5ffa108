"""Generates tests/golden/slam_regression.json: what the loop-closure driver (tbv_slam.ScanContextClosure) leaves behind on the seeded 1.4-lap
drive of tests/test_tbv_slam_cpu.py when the CPU oracle stands in for the device.  Like oracle_regression.json this is NOT a reference output:
it freezes oracle + host bookkeeping together, so that a change to either shows up as a diff (and next round's GPU run of the same driver can be
compared with it: same candidates, probabilities to 1e-6).
    python tests/golden/make_slam_regression.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def compute():
    import test_tbv_slam_cpu as T
    from tbv_slam_public_b200 import tbv_slam as TS
    g, gt, est = T.drive.__wrapped__()
    loop = TS.ScanContextClosure(g, T.OracleLoopDevice(), T._classifier(), TS.LoopClosureParams())
    loop.SearchAndAddConstraintBatched()
    recs = [[r.id_from, r.id_to, r.guess_nr, int(r.reg_ok), int(r.applied), round(r.probability, 9), [round(float(v), 9) for v in r.t_be],
             {k: round(float(v), 9) for k, v in sorted(r.quality.items())}] for r in loop.statistics]
    return {"keyframes": len(g), "records": recs, "constraints": [list(k) for k in sorted(loop.loop_constraints)]}


if __name__ == "__main__":
    data = compute()
    path = os.path.join(HERE, "slam_regression.json")
    json.dump(data, open(path, "w"), indent=0)
    print(path, os.path.getsize(path), "bytes", len(data["records"]), "records")
