"""Inference front end over a trained experiment: the part of the reference's
`Predictor` (/root/reference/predict.py:131-347) that sits either side of the segmentation path --

    results.txt  ->  hyper-parameters  ->  TextSegmenter.load_from_checkpoint  ->  tags per file  ->  sample ranges

Audio decoding / resampling / .wav writing (librosa, scipy.io.wavfile: predict.py:98-103, 316-345) and embedding
extraction (predict.py:35-86) are outside the path (SURVEY.md section 8 "out of scope"); `segment_ranges` gives the
(start, end) sample ranges the reference would cut, so a caller with an audio stack can finish the job.
"""
import os

from .EncoderDataset import AudioPortionDatasetInference, DevicePrefetcher
from .lightning_model import TextSegmenter
from .load_datasets_precomputed import load_dataset_for_inference
from .results_io import read_hyperparameters

# encoder-name prefix -> embedding width, first match wins (predict.py:184-216)
_ENCODER_DIMS = (("prosodic", 167), ("openl3_std", 1024), ("wav2vec_std", 1536), ("x-vector", 512), ("openl3", 512),
                 ("crepe_std", 512), ("crepe", 256), ("mfcc", 200), ("ecapa", 192), ("wav2vec", 768))


def encoder_embedding_dim(encoder, pca_reduce=False, pca_value=167):
    if pca_reduce:
        return pca_value
    for prefix, dim in _ENCODER_DIMS:
        if encoder.startswith(prefix):
            return dim
    raise ValueError("Encoder not recognised, use one of the three available options (x-vectors, ecapa or wav2vec)")


def segment_ranges(n_samples, segmentation, sr=16000, interval=1, adaptive=False):
    """Sample ranges cut at the predicted boundaries (predict.py:105-127): unit k ends at (k+1)*sr*interval samples
    (or (k+1)*(n_samples//100) with the adaptive 100-chunk scheme); a boundary at unit k closes a segment there.
    The fixed-interval scheme appends the remainder as a last segment and stops at the end of `segmentation`;
    the adaptive scheme does neither (and indexes past a short `segmentation` like the reference: IndexError)."""
    segs, prev = [], 0
    if adaptive:
        step = n_samples // 100
        for k, end in enumerate(range(step, n_samples + 1, step)):
            if segmentation[k]:
                segs.append((prev, end))
                prev = end
        return segs
    step = sr * int(interval)
    for k, end in enumerate(range(step, n_samples + 1, step)):
        if k >= len(segmentation):
            break
        if segmentation[k]:
            segs.append((prev, end))
            prev = end
    segs.append((prev, n_samples))
    return segs


class Predictor:
    """Same constructor arguments, attributes and error behaviour as predict.py:131-262 for the architectures on the
    path; `embedding_dim` may be given explicitly for encoders the name table does not know (multimodal concatenations)."""

    def __init__(self, hyperparameter_file, best_model_path, pca_reduce=False, pca_value=167, adaptive_uniform_interval=False,
                 uniform_interval=1, original_audio_extension=".mp3", threshold=0.5, sr=16000, embedding_dim=None,
                 device="cuda"):
        hp = read_hyperparameters(hyperparameter_file)
        self.encoder, self.architecture = hp.encoder, hp.architecture
        if embedding_dim is None:
            embedding_dim = encoder_embedding_dim(hp.encoder, pca_reduce, pca_value)
        if hp.architecture in ("SimpleBiLSTM", "BiLSTM"):
            bidirectional = True
        elif hp.architecture == "LSTM":
            bidirectional = False
        else:   # Transformer, BiLSTM-CRF, Transformer-CRF (predict.py:222-225)
            raise NotImplementedError()
        common = dict(architecture=hp.architecture, tagset_size=2,
                      embedding_dim=embedding_dim, hidden_dim=hp.hidden_units, bidirectional=bidirectional, lr=1e-3,
                      num_layers=hp.num_layers, dropout_in=0.0, dropout_out=0.0)
        try:
            self.model = TextSegmenter.load_from_checkpoint(best_model_path, loss_fn="BinaryCrossEntropy", threshold=threshold,
                                                            **common)
        except (KeyError, RuntimeError):
            # a CrossEntropy-trained head is [2, 2H] rather than [1, 2H] (predict.py:242-258; torch reports the
            # mismatch as RuntimeError, older Lightning as KeyError); the reference pins threshold 0.5 on this branch
            self.model = TextSegmenter.load_from_checkpoint(best_model_path, loss_fn="CrossEntropy", threshold=0.5, **common)
        self.model = self.model.to(device).eval()
        self.device = device
        self.adapt = bool(adaptive_uniform_interval)
        self.interval = uniform_interval
        self.ext = original_audio_extension
        self.th = threshold
        self.sr = sr

    def predict(self, embedding_folder, experiment_name, write_audio_segments=False, audio_directory=None, batch_size=1,
                num_gpus=1, verbose=False, add_overlap=1):
        """predict.py:264-347 up to the tags: one list of per-file tag lists per batch, in `os.listdir` order.
        The experiment directory is created (and must not exist) like the reference's; unlike it the working
        directory is left alone."""
        assert not os.path.exists(experiment_name), (
            "The name of this experiment has already be used: please change experiment name or delete all the existent "
            "results from {} folder to use this name".format(experiment_name))
        if write_audio_segments:
            raise NotImplementedError("audio decoding and .wav writing are outside the accelerated path; "
                                      "use segment_ranges() with your audio stack")
        os.makedirs(experiment_name)
        embeddings, self.file_names = load_dataset_for_inference(embedding_folder)
        if verbose:
            print(f"Segmenting the following files:\n{self.file_names}")
        data = AudioPortionDatasetInference(embeddings, encoder=self.encoder)
        if self.architecture == "SimpleBiLSTM":
            batch_size = 1
        batches = [data.collater([data[i] for i in range(lo, min(lo + batch_size, len(data)))])
                   for lo in range(0, len(data), batch_size)]
        if verbose:
            print("Test loader has: {} documents".format(len(data)))
        return list(self.model.predict_batches(DevicePrefetcher(batches, self.device)))
