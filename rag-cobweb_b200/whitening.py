"""PCA + ICA whitening model with the reference's interface (src/whitening/pca_ica.py).

`fit` stays on the CPU (sklearn PCA / FastICA, exactly the reference's recipe :54-76); `transform`
runs the two small GEMMs on the device through `cw_whiten` so whitened queries and documents
never round-trip through host numpy on their way into `ifit` / `predict`.  `save` / `load` use
the reference's pickle layout (:78-98).
"""
import pickle

import numpy as np
import torch

from . import _lib


class PCAICAWhiteningModel:
    def __init__(self, mean, pca_components, ica_unmixing, pca_explained_var, eps=1e-8):
        self.mean = np.asarray(mean)
        self.pca_components = np.asarray(pca_components)
        self.pca_explained_var = np.asarray(pca_explained_var)
        self.ica_unmixing = np.asarray(ica_unmixing)
        self.eps = eps
        self._dev = None

    def __repr__(self):
        return (f"{self.__class__.__name__}(\n  mean.shape={self.mean.shape},\n"
                f"  pca_components.shape={self.pca_components.shape},\n"
                f"  pca_explained_var.shape={self.pca_explained_var.shape},\n"
                f"  ica_unmixing.shape={self.ica_unmixing.shape},\n  eps={self.eps}\n)")

    def _device_params(self):
        if self._dev is None:
            _lib.require_cuda()
            f = lambda a: torch.as_tensor(np.ascontiguousarray(a, np.float32), device="cuda")  # noqa: E731
            scale = np.sqrt(self.pca_explained_var + self.eps).astype(np.float32)  # the reference's own expression (:42)
            self._dev = dict(mean=f(self.mean), pca=f(self.pca_components), scale=f(scale), ica=f(self.ica_unmixing))
        return self._dev

    def transform_device(self, x, is_ica=True):
        """x: [n, D_in] (numpy or tensor) -> torch CUDA tensor [n, K] (stays on the device)."""
        p = self._device_params()
        X = x.detach().to("cuda", torch.float32) if torch.is_tensor(x) else torch.as_tensor(
            np.ascontiguousarray(x, np.float32), device="cuda")
        X = X.reshape(1, -1) if X.dim() == 1 else X.contiguous()
        n, k = X.shape[0], p["pca"].shape[0]
        Y = torch.empty((n, k), dtype=torch.float32, device="cuda")
        tmp = torch.empty((n, k), dtype=torch.float32, device="cuda") if is_ica else None
        _lib.check(_lib.load().cw_whiten(X.data_ptr(), n, X.shape[1], p["mean"].data_ptr(), p["pca"].data_ptr(), k,
                                         p["scale"].data_ptr(), p["ica"].data_ptr() if is_ica else None,
                                         tmp.data_ptr() if is_ica else None, Y.data_ptr(), _lib.stream_ptr()), "cw_whiten")
        return Y

    def transform(self, x, is_ica=True):
        """PCAICAWhiteningModel.transform (pca_ica.py:30-51): numpy in, numpy out."""
        x = np.asarray(x)
        y = self.transform_device(x, is_ica).cpu().numpy()
        return y[0] if x.ndim == 1 else y

    @classmethod
    def fit(cls, X, pca_dim=256, eps=1e-8, ica_max_iter=5000, ica_tol=1e-3):
        """PCAICAWhiteningModel.fit (pca_ica.py:53-76), CPU / sklearn like the reference."""
        from sklearn.decomposition import PCA, FastICA
        mean = X.mean(axis=0)
        Xc = X - mean
        pca = PCA(n_components=pca_dim)
        X_pca = pca.fit_transform(Xc)
        components, explained_var = pca.components_, pca.explained_variance_
        X_pca_n = X_pca / np.sqrt(explained_var + eps)
        ica = FastICA(n_components=components.shape[0], whiten="unit-variance", max_iter=ica_max_iter, tol=ica_tol)
        ica.fit_transform(X_pca_n)
        return cls(mean, components, ica.components_, explained_var, eps)

    def save(self, filepath):
        with open(filepath, "wb") as f:
            pickle.dump({"mean": self.mean, "pca_components": self.pca_components,
                         "pca_explained_var": self.pca_explained_var, "ica_unmixing": self.ica_unmixing,
                         "eps": self.eps}, f)

    @classmethod
    def load(cls, filepath):
        with open(filepath, "rb") as f:
            d = pickle.load(f)
        return cls(mean=d["mean"], pca_components=d["pca_components"], pca_explained_var=d["pca_explained_var"],
                   ica_unmixing=d["ica_unmixing"], eps=d["eps"])
