"""Golden fixture for the loader mirror: builds a small synthetic dataset folder (two modality folders, labs pickle,
split json, timings pickle), runs the UNMODIFIED reference loader (/root/reference/utils/load_datasets_precomputed.py)
on it in standard-split mode, and stores the inputs and what the reference returned in tests/golden/loader_split.npz.
Run in the build container only (the reference does not exist on the GPU box)."""
import json
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from utils import load_datasets_precomputed as ref  # noqa: E402


def build(root, rng):
    names = [f"ep{i:02d}" for i in range(7)]
    lens = [5, 9, 3, 12, 7, 4, 6]
    os.makedirs(os.path.join(root, "text"))
    os.makedirs(os.path.join(root, "audio"))
    store, labs, times = {}, {}, {}
    for n, L in zip(names, lens):
        t = rng.standard_normal((L, 6)).astype(np.float32)
        a = rng.standard_normal((1, L, 4)).astype(np.float32)  # an extra leading axis: the loader squeezes it
        np.save(os.path.join(root, "text", n + ".npy"), t)
        np.save(os.path.join(root, "audio", n + ".npy"), a)
        lab = (rng.random(L) < 0.3).astype(int).tolist()
        lab[-1] = 1  # the loader must force it to 0
        labs[n] = lab
        times[n] = rng.random((L, 2)).astype(np.float32).tolist()
        store[n] = (t, a, lab, times[n])
    labs["ep02"] = labs["ep02"]
    with open(os.path.join(root, "labs_dict.pkl"), "wb") as f:
        pickle.dump(labs, f)
    with open(os.path.join(root, "times.pkl"), "wb") as f:
        pickle.dump(times, f)
    split = {"train": [n + ".npy" for n in names[:4]], "validation": [names[4] + ".npy"], "test": [n + ".npy" for n in names[5:]]}
    with open(os.path.join(root, "split.json"), "w") as f:
        json.dump(split, f)
    return names, store, split


def main():
    rng = np.random.default_rng(2024)
    out = {}
    with tempfile.TemporaryDirectory() as root:
        names, store, split = build(root, rng)
        for tag, timing in (("plain", None), ("timed", os.path.join(root, "times.pkl"))):
            res = ref.load_dataset_from_precomputed(os.path.join(root, "text") + "+" + os.path.join(root, "audio"),
                                                    os.path.join(root, "labs_dict.pkl"), split=os.path.join(root, "split.json"),
                                                    timing_info=timing)
            assert len(res) == 1 and len(res[0]) == 3
            for part, eps in zip(("train", "test", "validation"), res[0]):
                out[f"{tag}:{part}:names"] = np.array([e[2] for e in eps])
                for e in eps:
                    out[f"{tag}:{part}:{e[2]}:x"] = e[0].numpy()
                    out[f"{tag}:{part}:{e[2]}:y"] = np.array(e[1])
        for n in names:
            t, a, lab, tm = store[n]
            out[f"in:{n}:text"], out[f"in:{n}:audio"] = t, a
            out[f"in:{n}:labs"], out[f"in:{n}:times"] = np.array(lab), np.array(tm, dtype=np.float32)
        out["in:split"] = np.array(json.dumps(split))
    np.savez_compressed(os.path.join(HERE, "loader_split.npz"), **out)
    print("written", os.path.join(HERE, "loader_split.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
