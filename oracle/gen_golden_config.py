"""Generates tests/golden/common_paths.json from the UNMODIFIED reference config module
(/root/reference/hyper_params.py: defaults :50-80, get_common_path :3-48, data_dir rule :90-95).  The module is
imported from a scratch directory because it creates saved_logs/ and saved_models/ in the cwd (:87-88).

    python oracle/gen_golden_config.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    sys.path.insert(0, "/root/reference")
    import hyper_params as ref                                   # the reference's config module
    os.chdir(cwd)
    defaults = {k: v for k, v in ref.hyper_params.items() if k not in ("common_path", "log_file", "model_path", "data_dir")}
    cases = []
    for mt in ("bias_only", "MF", "MF_dot", "NeuMF", "deepconn", "deepconn++", "NARRE", "transnet", "transnet++", "HFT", "MPCN"):
        for over in ({}, {"dataset": "Beauty", "k_core": 0, "percent_reviews_to_keep": 50, "latent_size": 32, "word_embed_size": 300,
                          "lr": 0.01, "dropout": 0.3, "narre_num_words": 200}):
            hp = dict(defaults, model_type=mt, **over)
            if mt == "NARRE":
                hp["only_reviews"] = False                       # absent from the reference dict: KeyError otherwise (hyper_params.py:29)
            cases.append({"hyper_params": hp, "common_path": ref.get_common_path(hp)})
    out = {"defaults": defaults, "default_derived": {k: ref.hyper_params[k] for k in ("common_path", "log_file", "model_path", "data_dir")},
           "cases": cases}
    with open(os.path.join(ROOT, "tests", "golden", "common_paths.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(cases), "cases;", out["default_derived"])


if __name__ == "__main__":
    main()
