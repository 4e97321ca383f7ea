"""Generate the committed golden fixtures from the reference's own artefacts.

Run in the authoring container (needs /root/reference):
    python tests/golden/make_golden.py

Outputs (committed, tiny):
  tests/golden/svcdata.npz        -- the reference's SVCdata data set (R-pkg/data/SVCdata.rda,
                                     500 rows) + the vignette / testthat sub-sample indices,
                                     reproduced with the R RNG replica (oracle/r_rng.py).
  tests/golden/vignette_golden.json -- numbers printed in the knitted vignette
                                     examples/01_Introduction.html (source
                                     R-pkg/vignettes/Introduction.Rmd:109-165).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.rdata import read_rda, as_python  # noqa: E402
from oracle.r_rng import RRandom  # noqa: E402

REF = "/root/reference"


def main():
    d = as_python(read_rda(os.path.join(REF, "R-pkg/data/SVCdata.rda"))["SVCdata"])
    r = RRandom(123)
    id_train = np.array(sorted(r.sample(500, 100)))          # Introduction.Rmd:113-115
    s_id = r.sample(500, 1)[0]                               # later chunk, printed "159"
    id_test50 = np.array(RRandom(123).sample(500, 50))        # test_SVC_mle.R:6-7 (unsorted)
    np.savez(os.path.join(HERE, "svcdata.npz"),
             y=d["y"], X=d["X"], beta=d["beta"], eps=d["eps"], locs=d["locs"],
             true_var=d["true_pars"]["var"], true_scale=d["true_pars"]["scale"],
             true_mean=d["true_pars"]["mean"], true_nu=d["true_pars"]["nu"],
             id_train=id_train, id_test50=id_test50, s_id=s_id)
    golden = {
        "source": "examples/01_Introduction.html (knitted R-pkg/vignettes/Introduction.Rmd)",
        "fit": "SVC_mle(y ~ x, data=df_train, locs=df_train$locs): exp cov, non-profile, n=100",
        "id_train_head": [4, 7, 13, 14, 16, 23],
        "s_id": 159,
        "lm_coef": [0.7703167, 2.6885343],
        "coef": [0.6490864, 2.5468056],
        "coef_se": [0.2838, 0.3499],
        "cov_par": [1.68358, 0.25786, 1.31196, 0.51061, 0.26822],
        "cov_par_se": [1.44902, 0.15979, 1.16981, 0.31030, 0.05479],
        "logLik": -106.7,
        "logLik_2dp": -106.70,
        "BIC": 231.82,
        "cAIC": 284.65,
        "fn_evals": 76,
        "convergence": 0,
        "residual_summary": [-1.25497, -0.26972, 0.01098, 0.30372, 0.98823],
        "residual_se": 0.4307,
        "fitted_head": {
            "ids": [4, 7, 13, 14, 16, 23],
            "SVC_1": [-0.3353062, -0.3390177, -0.2881019, -0.2775242, -0.2681113, -0.2584885],
            "SVC_2": [0.5712240, 0.5691786, 0.6498610, 0.6628253, 0.6705916, 0.8309101],
            "y_pred": [-3.1805285, 1.6638651, 0.2224001, 0.2605299, -3.5324973, 6.9425422],
            "loc_1": [0.0821552, 0.1776581, 0.4205953, 0.4555650, 0.4766363, 0.5847849],
        },
        "str_head": {"y": [-1.488, 3.901, 0.197, -3.326, 2.722],
                     "locs": [0.00465, 0.00625, 0.06301, 0.08216, 0.0943]},
    }
    with open(os.path.join(HERE, "vignette_golden.json"), "w") as f:
        json.dump(golden, f, indent=1)
    print("wrote fixtures:", id_train[:6], s_id)


if __name__ == "__main__":
    main()
