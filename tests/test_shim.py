"""The stand-in torchreid package (ieee_b200/shim) and the drop-in signatures: same names, argument names and
defaults as the reference's functions (torchreid/metrics/distance.py:6, rank.py:246, utils/rerank.py:31,
utils/reidtools.py:18).  The expected signatures below were read off the reference sources; where
/root/reference is present (build container) they are re-checked against the files themselves."""
import ast
import inspect
import os
import subprocess
import sys

import pytest

REF = "/root/reference/torchreid"
EXPECTED = {
    "compute_distance_matrix": (["input1", "input2", "metric"], ["euclidean"]),
    "evaluate_rank": (["distmat", "q_pids", "g_pids", "q_camids", "g_camids", "max_rank", "use_metric_cuhk03", "use_cython"],
                      [20, False, True]),
    "re_ranking": (["q_g_dist", "q_q_dist", "g_g_dist", "k1", "k2", "lambda_value"], [20, 6, 0.3]),
    "visualize_ranked_results": (["distmat", "dataset", "data_type", "width", "height", "save_dir", "topk"],
                                 [128, 256, "", 10]),
}
WHERE = {"compute_distance_matrix": "metrics/distance.py", "evaluate_rank": "metrics/rank.py",
         "re_ranking": "utils/rerank.py", "visualize_ranked_results": "utils/reidtools.py"}


def ours():
    from ieee_b200.metrics import compute_distance_matrix, evaluate_rank
    from ieee_b200.utils import re_ranking, visualize_ranked_results
    return {"compute_distance_matrix": compute_distance_matrix, "evaluate_rank": evaluate_rank, "re_ranking": re_ranking,
            "visualize_ranked_results": visualize_ranked_results}


@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_drop_in_signatures(name):
    names, defaults = EXPECTED[name]
    params = list(inspect.signature(ours()[name]).parameters.values())
    assert [p.name for p in params[:len(names)]] == names                      # same positional order and names
    got_defaults = [p.default for p in params[:len(names)] if p.default is not inspect.Parameter.empty]
    assert got_defaults == defaults
    for extra in params[len(names):]:                                          # extensions must be optional
        assert extra.default is not inspect.Parameter.empty


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("name", sorted(EXPECTED))
def test_expected_signatures_are_the_references(name):
    tree = ast.parse(open(os.path.join(REF, WHERE[name])).read())
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == name)
    assert [a.arg for a in fn.args.args] == EXPECTED[name][0]
    assert [ast.literal_eval(d) for d in fn.args.defaults] == EXPECTED[name][1]


def test_shim_package_resolves_torchreid_names():
    import ieee_b200.shim as shim
    code = (
        "import torchreid, ieee_b200\n"
        "import torchreid.metrics as m, torchreid.utils as u\n"
        "from torchreid.metrics.distance import compute_distance_matrix as c1\n"
        "from torchreid.metrics.rank import evaluate_rank as e1\n"
        "from torchreid.utils.rerank import re_ranking as r1\n"
        "from torchreid.utils.reidtools import visualize_ranked_results as v1\n"
        "assert torchreid.__file__.startswith(%r), torchreid.__file__\n"
        "assert m.compute_distance_matrix is c1 is ieee_b200.metrics.compute_distance_matrix\n"
        "assert m.evaluate_rank is e1 is ieee_b200.metrics.evaluate_rank\n"
        "assert u.re_ranking is r1 is ieee_b200.utils.re_ranking\n"
        "assert u.visualize_ranked_results is v1\n"
        "print('ok')\n" % shim.PATH)
    env = dict(os.environ)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = os.pathsep.join([shim.PATH, root])
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_result_lines_parse_with_the_references_log_parser():
    """tools/parse_test_res.py:70-74 greps these five patterns out of test.log; the device-resident engine prints
    through ieee_b200.engine.result_lines, so its logs keep parsing."""
    import re

    import numpy as np
    from ieee_b200.engine import result_lines
    cmc = np.linspace(0.5, 0.99, 20).astype(np.float32)
    lines = result_lines(cmc, 0.61049, ranks=[1, 5, 10, 20])
    assert lines[:3] == ["** Results **", "mAP: 61.05%", "CMC curve"] and lines[-1] == "\n"
    patterns = {"mAP": r'mAP: ([\.\deE+-]+)%', "r1": r'Rank-1  : ([\.\deE+-]+)%', "r5": r'Rank-5  : ([\.\deE+-]+)%',
                "r10": r'Rank-10 : ([\.\deE+-]+)%', "r20": r'Rank-20 : ([\.\deE+-]+)%'}
    found = {}
    for line in lines:
        for key, pat in patterns.items():
            m = re.compile(pat).search(line.strip())
            if m:
                found[key] = float(m.group(1))
    assert found == {"mAP": 61.05, "r1": 50.0, "r5": round(float(cmc[4]) * 100, 2), "r10": round(float(cmc[9]) * 100, 2),
                     "r20": 99.0}
    if os.path.isfile("/root/reference/tools/parse_test_res.py"):            # the patterns above are the reference's
        src = open("/root/reference/tools/parse_test_res.py").read()
        for pat in patterns.values():
            assert "r'%s'" % pat in src
    assert result_lines(cmc[:3], 0.5, ranks=[1, 5]) == ["** Results **", "mAP: 50.00%", "CMC curve", "Rank-1  : 50.00%", "\n"]
