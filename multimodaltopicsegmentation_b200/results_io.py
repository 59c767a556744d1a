"""Result files either side of the segmentation path.

The reference's driver script leaves four artefacts in the experiment directory and its inference
front end reads one of them back; this module writes and parses the same formats so an experiment
directory produced with this package is readable by the reference's tools and vice versa:

  results.txt        hyper-parameter + mean-score summary, one "\\n<line>\\n" record per line
                     (written at /root/reference/train_fit.py:583-624, parsed at predict.py:168-177)
  logs               per-fold score lines appended after every fold (train_fit.py:400-412)
  all_results.json   {test file id: that file's metric dict}  (train_fit.py:431-447, 460-462)
  all_scores.json    {test file id: per-sentence boundary scores}  (train_fit.py:449-451, 464-465)

Host-side text only; nothing here touches the device.
"""
import json
import os
from dataclasses import dataclass

# metric == 'B' / 'scaiano' relabel Pk/WD as precision/recall (train_fit.py:575-578)
_BOUNDARY_LABELS = {"Pk": "Precision", "WD": "Recall", "F1": "F1"}
_WINDOW_LABELS = {"Pk": "Pk", "WD": "WD", "F1": "F1"}


def _labels(metric):
    return _BOUNDARY_LABELS if str(metric).lower() in ("b", "scaiano") else _WINDOW_LABELS


@dataclass
class Hyperparameters:
    """What predict.py:168-177 recovers from results.txt."""
    encoder: str
    architecture: str
    hidden_units: int
    num_layers: int


def summary_lines(experiment_name, encoder, architecture, batch_size, hidden_units, dropout_in, dropout_out, num_layers,
                  optimizer, best_results, confidence=None, metric="Pk", zero_shot_labels=None):
    """The record list of results.txt.  `best_results` holds 'Pk', 'F1', 'WD' (and 'B' when metric == 'B');
    `confidence` (same keys) adds the bootstrap half-widths the cross-validation branch prints
    (train_fit.py:607-619); None gives the held-out-test form (train_fit.py:583-599)."""
    lab = _labels(metric)
    lines = [f"Results for experiment {experiment_name} with following parameters:",
             f"Sentence encoder: {encoder}",
             f"Neural architecture: {architecture}",
             f"Batch size: {batch_size}",
             f"Hidden units: {hidden_units}",
             f"Dropout in: {dropout_in}",
             f"Dropout out: {dropout_out}",
             f"Number of layers: {num_layers}",
             f"Optimizer: {optimizer}"]

    def mean(name, key):
        s = f"Mean {name} obtained is {best_results[key]}"
        if confidence is not None:
            s += f" with a 95% confidence interval of +- {confidence[key]}"
        return s

    lines += [mean(lab["Pk"], "Pk"), mean("F1", "F1"), mean(lab["WD"], "WD")]
    if str(metric).lower() == "b":
        lines.append(mean("Boundary Similarity", "B"))
    if zero_shot_labels is not None:
        lines.append("Labels: " + str(zero_shot_labels))
    return lines


def write_results_txt(directory, lines, name="results.txt"):
    """Every record is preceded and followed by a newline (train_fit.py:622-624)."""
    path = os.path.join(directory, name)
    with open(path, "w") as f:
        for line in lines:
            f.write("\n" + line + "\n")
    return path


def read_hyperparameters(path):
    """The parse of predict.py:168-177: whitespace-split fields at fixed positions of four record kinds.
    A file that lacks any of them is an error here (the reference fails later with an unbound name)."""
    found = {}
    with open(path) as f:
        for line in f:
            if line.startswith("Sentence encoder"):
                found["encoder"] = line.split()[2]
            elif line.startswith("Neural architecture"):
                found["architecture"] = line.split()[2]
            elif line.startswith("Hidden units"):
                found["hidden_units"] = int(line.split()[2])
            elif line.startswith("Number of layers"):
                found["num_layers"] = int(line.split()[3])
    missing = [k for k in ("encoder", "architecture", "hidden_units", "num_layers") if k not in found]
    if missing:
        raise ValueError(f"{path}: no record for {', '.join(missing)}")
    return Hyperparameters(**found)


def fold_keys(metric):
    """Which entries of the logged test dict hold Pk / WD / F1 (and B): the tuned metric travels as
    'test_loss', the others under their own names (train_fit.py:375-397)."""
    m = str(metric).lower()
    if metric == "F1":
        return {"Pk": "Pk_loss", "WD": "WD_loss", "F1": "test_loss"}
    if metric == "WD":
        return {"Pk": "Pk_loss", "WD": "test_loss", "F1": "F1_loss"}
    if m == "b":
        return {"Pk": "b_precision", "WD": "b_recall", "F1": "b_f1", "B": "test_loss"}
    if m == "scaiano":
        return {"Pk": "b_precision", "WD": "b_recall", "F1": "test_loss"}
    return {"Pk": "test_loss", "WD": "WD_loss", "F1": "F1_loss"}


def fold_metrics(fold_results, metric="Pk"):
    """{'Pk','WD','F1'[,'B']} of one fold out of the logged dict (train_fit.py:421-437)."""
    return {name: fold_results[key] for name, key in fold_keys(metric).items()}


def append_fold_log(directory, fold, fold_results, metric="Pk", name="logs"):
    """One fold's block of the `logs` file (train_fit.py:399-412)."""
    m = str(metric).lower()
    v = fold_metrics(fold_results, metric)
    with open(os.path.join(directory, name), "a") as f:
        f.write(f"Results for fold number {fold}\n")
        if m in ("b", "scaiano"):
            f.write(f"B_precision score: {v['Pk']}\n")
            f.write(f"B_recall score: {v['WD']}\n")
            f.write(f"B_F1 score: {v['F1']}\n")
            if m == "b":
                f.write(f"B Similarity score: {v['B']}\n")
        else:
            f.write(f"PK score: {v['Pk']}\n")
            f.write(f"WD score: {v['WD']}\n")
            f.write(f"F1 score: {v['F1']}\n")


def read_fold_logs(path):
    """[(fold, {label: value})] back from a `logs` file."""
    folds = []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith("Results for fold number"):
                folds.append((int(line.split()[-1]), {}))
            elif " score: " in line and folds:
                label, value = line.split(" score: ")
                folds[-1][1][label] = float(value)
    return folds


def collect_per_file(test_files, per_file_results, per_file_scores=None, metric="Pk", all_results=None, all_scores=None):
    """Fold one test split into the two per-file maps.  `test_files[i][2]` is the file id
    (the dataset tuples are (embeddings, labels, name), load_datasets_precomputed); the per-file metric
    dict has its 'test_loss' entry renamed to the tuned metric's name (train_fit.py:436-441)."""
    all_results = {} if all_results is None else all_results
    all_scores = {} if all_scores is None else all_scores
    for i, item in enumerate(test_files):
        entry = dict(per_file_results[i])
        if "test_loss" in entry:
            entry[metric] = entry.pop("test_loss")
        all_results[item[2]] = {k: _plain(v) for k, v in entry.items()}
        if per_file_scores is not None:
            all_scores[item[2]] = _plain(per_file_scores[i])
    return all_results, all_scores


def _plain(v):
    """json-serialisable copy of a tensor / array / scalar."""
    if hasattr(v, "detach"):
        v = v.detach().cpu()
    if hasattr(v, "tolist"):
        return v.tolist()
    return v


def write_per_file_json(directory, all_results=None, all_scores=None):
    if all_results is not None:
        with open(os.path.join(directory, "all_results.json"), "w") as f:
            json.dump(all_results, f)
    if all_scores is not None:
        with open(os.path.join(directory, "all_scores.json"), "w") as f:
            json.dump(all_scores, f)


def read_per_file_json(directory):
    out = []
    for name in ("all_results.json", "all_scores.json"):
        path = os.path.join(directory, name)
        if os.path.exists(path):
            with open(path) as f:
                out.append(json.load(f))
        else:
            out.append(None)
    return tuple(out)
