"""k-nearest-neighbour evaluation on the B200.

Mirror of what the reference does with scikit-learn on the host every 10 epochs: ``compute_knn``
(/root/reference/src/train_utils/knn.py:22-42) fits ``KNeighborsClassifier()`` (k = 5, Euclidean, uniform weights) on the
training embeddings and ``eval_pretrained_model`` (/root/reference/src/train_utils/eval_functions.py:65-97) calls
``estimator.predict`` on the validation embeddings, after moving every batch to the CPU.  Here the embeddings stay in HBM
and two hand-written kernels (tiled direct-difference distances, warp-per-query selection + vote) produce the labels.

``KNNEstimator`` has sklearn's ``fit`` / ``predict`` / ``kneighbors`` shape so that ``eval_pretrained_model`` can use it
unchanged (it accepts and returns numpy arrays as well as CUDA tensors); ``compute_knn`` has the reference function's
signature.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _cabi


class KNNEstimator:
    def __init__(self, n_neighbors: int = 5, device: Optional[torch.device] = None):
        if not 1 <= n_neighbors <= 16:
            raise ValueError("n_neighbors must be in [1, 16]")
        self.k = n_neighbors
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._x: Optional[torch.Tensor] = None
        self._y: Optional[torch.Tensor] = None
        self.classes_: Optional[np.ndarray] = None

    def _to_dev(self, a, dtype):
        if isinstance(a, np.ndarray):
            a = torch.from_numpy(np.ascontiguousarray(a))
        if not isinstance(a, torch.Tensor):
            raise TypeError(f"expected a numpy array or a torch tensor, got {type(a)}")
        return a.to(self.device, dtype).contiguous()

    def fit(self, X, y) -> "KNNEstimator":
        x = self._to_dev(X, torch.float32)
        yy = self._to_dev(y, torch.int64).flatten()
        if x.dim() != 2 or x.shape[0] != yy.shape[0]:
            raise ValueError("X must be [n, dim] and y [n]")
        if x.shape[0] < self.k:
            raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = {self.k}, n_samples_fit = {x.shape[0]}")
        classes, inv = torch.unique(yy, return_inverse=True)          # sklearn: labels -> indices into classes_
        if classes.numel() > 32:
            raise ValueError("focal_b200.knn supports up to 32 classes")
        self._x, self._y = x, inv.to(torch.int32).contiguous()
        self.classes_ = classes.cpu().numpy()
        return self

    def _run(self, Q, want_neighbours: bool):
        if self._x is None:
            raise RuntimeError("fit() first")
        was_numpy = isinstance(Q, np.ndarray)
        q = self._to_dev(Q, torch.float32)
        if q.dim() != 2 or q.shape[1] != self._x.shape[1]:
            raise ValueError("query must be [n, dim] with the training dimension")
        nq, nt, dim = q.shape[0], self._x.shape[0], q.shape[1]
        lib = _cabi.load()
        d2 = torch.empty((nq, nt), dtype=torch.float32, device=self.device)
        out = torch.empty(nq, dtype=torch.int32, device=self.device)
        nbr = torch.empty((nq, self.k), dtype=torch.int32, device=self.device) if want_neighbours else None
        with torch.cuda.device(self.device):
            rc = lib.focal_b200_knn_predict(
                C.c_void_p(self._x.data_ptr()), C.c_void_p(self._y.data_ptr()), nt, C.c_void_p(q.data_ptr()), nq, dim,
                self.k, int(len(self.classes_)), C.c_void_p(d2.data_ptr()), C.c_void_p(out.data_ptr()),
                C.c_void_p(nbr.data_ptr()) if nbr is not None else None,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        _cabi.check(rc, "focal_b200_knn_predict")
        return out, nbr, d2, was_numpy

    def predict(self, Q):
        out, _, _, was_numpy = self._run(Q, False)
        labels = torch.from_numpy(self.classes_).to(self.device)[out.long()]
        return labels.cpu().numpy() if was_numpy else labels

    def kneighbors(self, Q):
        """(distances [n, k], indices [n, k]) like sklearn's, nearest first."""
        _, nbr, d2, was_numpy = self._run(Q, True)
        dist = torch.gather(d2, 1, nbr.long()).sqrt()
        return (dist.cpu().numpy(), nbr.cpu().numpy().astype(np.int64)) if was_numpy else (dist, nbr.long())


def compute_knn(args, classifier, augmenter, data_loader_train) -> KNNEstimator:
    """Same signature and role as the reference's ``compute_knn`` (knn.py:22-42); the embeddings stay on the device."""
    classifier.eval()
    feats, labels = [], []
    with torch.no_grad():
        for time_loc_inputs, y in data_loader_train:
            aug_freq_loc_inputs, _ = augmenter.forward("no", time_loc_inputs, y)
            mod_features = classifier(aug_freq_loc_inputs, class_head=False, proj_head=False)
            feats.append(torch.cat([mod_features[m] for m in args.dataset_config["modality_names"]], dim=1).float())
            yy = y.argmax(dim=1) if y.dim() > 1 else y
            labels.append(yy.to(feats[-1].device))
    return KNNEstimator(5, feats[0].device).fit(torch.cat(feats), torch.cat(labels))
