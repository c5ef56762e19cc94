"""PCA — drop-in for `imagescry/models/decomposition.py`, projection computed on tcgen05.

`PCA.forward` / `PCA.transform` (`decomposition.py:79-91,150-165`) run the fused sm_100a projection
kernel (`isx_l2norm_project` with `normalize=0`); `project_feature_map` exposes the fully fused
L2-normalise (+pool) + projection of a backbone feature map.  `fit` (`:94-148`) takes its means
and covariance from the sm_100a moment kernels (`isx_pca_moments`, a tcgen05 SYRK).

The reference derives from `LightningModule`; lightning is not part of this image, so the base is
`torch.nn.Module` with the same hyper-parameter surface (`hparams`, `save_hyperparameters`).
"""

from __future__ import annotations

import torch
from jaxtyping import Float, jaxtyped
from torch import Tensor, nn

from imagescry_b200 import _lib
from imagescry_b200.typechecking import typechecker


class _HParams(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class HParamsModule(nn.Module):
    """nn.Module with the slice of LightningModule's hyper-parameter API the reference uses."""

    def __init__(self) -> None:
        super().__init__()
        self._hparams = _HParams()

    def save_hyperparameters(self, hp: dict | None = None) -> None:
        if hp:
            self._hparams.update(dict(hp))

    @property
    def hparams(self) -> _HParams:
        return self._hparams


def select_num_components(
    explained_variance: Tensor, min_num_components: int, max_num_components: int | None, min_explained_variance: float
) -> int:
    """The reference's selection rule (`decomposition.py:128-137`): the fewest components whose
    cumulative explained variance reaches `min_explained_variance`, clamped to [min, max]."""
    cumulative = torch.cumsum(explained_variance, dim=0)
    need = int(torch.sum(cumulative < min_explained_variance).item() + 1)
    num_components = max(min_num_components, need)
    if max_num_components is not None:
        num_components = min(max_num_components, num_components)
    return num_components


class PCA(HParamsModule):
    """Principal component analysis: linear projection to a lower dimensional space via SVD."""

    def __init__(
        self,
        *,
        min_num_components: int = 1,
        max_num_components: int | None = None,
        min_explained_variance: float = 0.0,
        num_features: int = 0,
        num_components: int = 0,
    ) -> None:
        super().__init__()
        if min_num_components < 1:
            raise ValueError(f"min_num_components must be at least 1, got {min_num_components}")
        if max_num_components is not None and max_num_components < min_num_components:
            raise ValueError(f"max_num_components must be at least {min_num_components}, got {max_num_components}")
        if min_explained_variance < 0.0 or min_explained_variance > 1.0:
            raise ValueError(f"min_explained_variance must be between 0.0 and 1.0, got {min_explained_variance}")

        self.min_num_components = min_num_components
        self.max_num_components = max_num_components
        self.min_explained_variance = min_explained_variance
        self.save_hyperparameters({
            "min_num_components": min_num_components,
            "max_num_components": max_num_components,
            "min_explained_variance": min_explained_variance,
        })

        self._fitted = nn.Parameter(torch.tensor(False), requires_grad=False)
        self._num_features = nn.Parameter(torch.tensor(0), requires_grad=False)
        self._num_components = nn.Parameter(torch.tensor(0), requires_grad=False)
        self.feature_means = nn.Parameter(torch.empty((1, num_features)), requires_grad=False)
        self.explained_variance = nn.Parameter(torch.empty((num_features,)), requires_grad=False)
        self.component_vectors = nn.Parameter(torch.empty((num_features, num_components)), requires_grad=False)
        self.__dict__["_packed"] = None
        self.__dict__["_packed_key"] = None
        self.__dict__["_fitted_key"] = None

    def __setattr__(self, name: str, value) -> None:
        # re-assigned weights invalidate the packed operand (a new tensor may reuse a freed address)
        if name in ("component_vectors", "feature_means", "_fitted"):
            self.__dict__["_packed"] = None
            self.__dict__["_packed_key"] = None
            self.__dict__["_fitted_key"] = None
        super().__setattr__(name, value)

    def _load_from_state_dict(self, *args, **kwargs):
        self.__dict__["_packed"] = None
        self.__dict__["_packed_key"] = None
        self.__dict__["_fitted_key"] = None
        return super()._load_from_state_dict(*args, **kwargs)

    def __repr__(self) -> str:
        num_features = self.num_features if self.fitted else "not fitted"
        num_components = self.num_components if self.fitted else "not fitted"
        return f"{self.__class__.__name__}(num_features={num_features}, num_components={num_components})"

    # ------------------------------------------------------------------ packed weights
    def _param_key(self) -> tuple:
        """Identity of the fitted parameters (storage, version counter, shape) — no device access."""
        cv, fm = self.component_vectors, self.feature_means

        def version(t: Tensor) -> int:
            try:
                return t._version
            except RuntimeError:  # inference tensors (fitted under torch.inference_mode) have no counter
                return -1

        return (cv.data_ptr(), fm.data_ptr(), version(cv), version(fm), tuple(cv.shape), tuple(cv.stride()), str(cv.device))

    def packed_weights(self) -> Tensor:
        """bf16 hi/lo split of `component_vectors` plus the bias -(means . components), in the layout
        the projection kernel's TMA descriptors read.  Rebuilt when the parameters change."""
        cv = self.component_vectors
        _lib.require_cuda(cv, "PCA parameters")
        fm = self.feature_means
        key = self._param_key()
        if self._packed is None or self._packed_key != key:
            lib = _lib.load()
            F, k = cv.shape
            nbytes = lib.isx_project_packed_bytes(F, k)
            if nbytes == 0:
                raise ValueError(f"PCA projection kernel supports 1..256 components, got {k} (features={F})")
            packed = torch.empty(nbytes, dtype=torch.uint8, device=cv.device)
            means = fm.detach().reshape(-1).contiguous().float()
            comps = cv.detach().float()
            with _lib.on_device(means, comps) as stream:
                rc = lib.isx_project_pack(
                    means.data_ptr(), comps.data_ptr(), F, k, comps.stride(0), comps.stride(1), packed.data_ptr(),
                    nbytes, stream,
                )
            _lib.check(rc, "isx_project_pack")
            self._packed, self._packed_key = packed, key
        return self._packed

    # ------------------------------------------------------------------ projection
    @jaxtyped(typechecker=typechecker)
    def forward(
        self, x: Float[Tensor, "num_samples {self.num_features}"]
    ) -> Float[Tensor, "num_samples {self.num_components}"]:
        """Project the input data: `(x - feature_means) @ component_vectors` (`decomposition.py:91`)."""
        _lib.require_cuda(x, "x")
        n, F = x.shape
        k = self.num_components
        out = torch.empty((n, k), dtype=torch.float32, device=x.device)
        if n == 0:
            return out
        xf = x.float().contiguous()
        lib = _lib.load()
        packed = self.packed_weights()
        with _lib.on_device(xf, packed) as stream:
            rc = lib.isx_l2norm_project(xf.data_ptr(), n, F, 1, 1, 0, 0, packed.data_ptr(), k, out.data_ptr(), None, 0, stream)
        _lib.check(rc, "isx_l2norm_project")
        return out

    def project_feature_map(
        self, fmap: Float[Tensor, "B E H W"], *, pool: str | None = None, precision: str = "exact"
    ) -> Tensor:
        """Fused stage 2 on a backbone feature map: L2-normalise every cell over the channels
        (`embedding.py:74`), optionally mean-pool the cells, project (`decomposition.py:91`).

        pool=None  → B×k×h×w with NHWC strides, the exact tensor `pipelines.py:82-84` returns.
        pool="mean" → B×k.
        precision="exact" → three-pass bf16 split (fp32-class products, ~1e-6 of a row's norm);
        precision="fp16"  → one fp16 tensor pass (~1e-5 of a row's norm, |x| ≤ 65504): the kernel
                            then runs at the HBM roofline instead of the tensor one.
        """
        # `fitted`, `num_features` and `num_components` are scalar parameters (decomposition.py keeps
        # them in the state dict); reading them is a device synchronisation each, and three per call
        # left the GPU idle for ~16 % of a pooled-projection step.  The fitted flag is therefore
        # re-read only when the parameters changed, and the sizes come from the weight matrix
        # (`fit` stores exactly num_features x num_components of it, decomposition.py:139-146).
        key = self._param_key()
        if getattr(self, "_fitted_key", None) != key:
            if not self.fitted:
                raise RuntimeError("PCA model not fitted")
            self._fitted_key = key
        _lib.require_cuda(fmap, "fmap")
        if pool not in (None, "mean"):
            raise ValueError(f"Invalid pool: {pool}")
        if precision not in ("exact", "fp16"):
            raise ValueError(f"Invalid precision: {precision}")
        B, E, h, w = fmap.shape
        num_features, k = self.component_vectors.shape
        if E != num_features:
            raise ValueError(f"feature map has {E} channels, PCA was fitted on {num_features}")
        f = fmap.float().contiguous()
        lib = _lib.load()
        if pool is None:
            out = torch.empty((B, h, w, k), dtype=torch.float32, device=f.device)
            ws, ws_bytes = None, 0
        else:
            out = torch.empty((B, k), dtype=torch.float32, device=f.device)
            ws_bytes = lib.isx_l2norm_project_workspace_bytes(B, E, h, w, k, 1)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=f.device)
        if B > 0:
            entry = lib.isx_l2norm_project if precision == "exact" else lib.isx_l2norm_project_fp16
            packed = self.packed_weights()
            with _lib.on_device(f, packed) as stream:
                rc = entry(
                    f.data_ptr(), B, E, h, w, int(pool is not None), 1, packed.data_ptr(), k,
                    out.data_ptr(), None if ws is None else ws.data_ptr(), ws_bytes, stream,
                )
            _lib.check(rc, "isx_l2norm_project")
        return out.permute(0, 3, 1, 2) if pool is None else out

    # ------------------------------------------------------------------ fit
    def _moments(self, x: Tensor) -> tuple[Tensor, Tensor]:
        """Feature means (F,) and covariance (F×F) of a CUDA matrix through `isx_pca_moments`."""
        xf = x.float().contiguous()
        n, F = xf.shape
        lib = _lib.load()
        mean = torch.empty(F, dtype=torch.float32, device=xf.device)
        cov = torch.empty((F, F), dtype=torch.float32, device=xf.device)
        ws_bytes = int(lib.isx_pca_moments_workspace_bytes(n, F))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xf.device)
        with _lib.on_device(xf) as stream:
            rc = lib.isx_pca_moments(xf.data_ptr(), n, F, mean.data_ptr(), cov.data_ptr(), ws.data_ptr(), ws_bytes, stream)
        _lib.check(rc, "isx_pca_moments")
        return mean, cov

    @jaxtyped(typechecker=typechecker)
    def fit(self, x: Float[Tensor, "num_samples num_features"]) -> "PCA":
        """Fit (`decomposition.py:94-148`): same component-selection rule (min/max components,
        minimum explained variance) and the same fitted tensors as the reference.

        Means and the F×F covariance of the fp32-centred data come from the sm_100a moment kernels
        (`isx_pca_moments`: fp64 column means, tcgen05 SYRK with a bf16 hi/lo split); the covariance's
        eigen-pairs are the reference's `vt` rows and `s²/(n-1)` (:122-125) without the full SVD's n×n
        `U`.  The F×F `eigh` (fp64) is a small dense library call.  Signs of the component vectors
        are as arbitrary as the reference's own.  CUDA tensors only: there is no CPU path.

        Numerical note: eigenvalues come from the covariance (condition number squared relative to
        the SVD of the data), so explained-variance ratios below ~1e-7 of the total are noise; the
        reference's SVD resolves ~1e-14.  Component selection differs only if `min_explained_variance`
        falls within that distance of a cumulative-variance step."""
        _lib.require_cuda(x, "x")
        num_samples, num_features = x.shape
        if num_samples < 2:
            raise ValueError(f"num_samples must be at least 2, got {num_samples}")
        self._num_features = nn.Parameter(torch.tensor(num_features), requires_grad=False)
        mean, cov = self._moments(x)
        self.feature_means = nn.Parameter(mean.reshape(1, -1), requires_grad=False)
        evals, evecs = torch.linalg.eigh(cov.double())  # ascending; 6.5 MB problem at F = 1280
        rank = min(num_samples, num_features)  # the SVD's `s` has min(n, F) entries (:122-125)
        eigenvalues = evals.flip(0).clamp_min(0.0).float()[:rank]
        vt = evecs.flip(1).T.float()  # rows = principal directions, like the SVD's vt
        total_variance = torch.sum(eigenvalues)
        self.explained_variance = nn.Parameter(eigenvalues / total_variance, requires_grad=False)
        num_components = select_num_components(
            self.explained_variance, self.min_num_components, self.max_num_components, self.min_explained_variance
        )
        self._num_components = nn.Parameter(torch.tensor(num_components), requires_grad=False)
        self.component_vectors = nn.Parameter(vt[:num_components, :].T, requires_grad=False)
        self._fitted = nn.Parameter(torch.tensor(True), requires_grad=False)
        self.hparams.update({"num_features": num_features, "num_components": num_components})
        return self

    @jaxtyped(typechecker=typechecker)
    def transform(self, x: Float[Tensor, "num_samples num_features"]) -> Float[Tensor, "num_samples num_components"]:
        """Project the input data (`decomposition.py:150-165`).

        Raises:
            RuntimeError: If the PCA model is not fitted.
        """
        if not self.fitted:
            raise RuntimeError("PCA model not fitted")
        return self(x)

    @property
    def fitted(self) -> bool:
        return bool(self._fitted.item())

    @property
    def num_features(self) -> int:
        return int(self._num_features.item())

    @property
    def num_components(self) -> int:
        return int(self._num_components.item())
