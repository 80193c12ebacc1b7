"""ctypes binding of ``libflexs_b200.so`` (the C ABI declared in include/flexs_b200.h).

This is the stub a maintainer of the reference would add next to
``flexs/baselines/models/keras_model.py``.  There is NO fallback: if the
shared library is missing, or no CUDA device is present, every compute call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_uint8, c_uint64, c_void_p
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

# FLEXS_B200_LIB: load another build of the same library (A/B timing of kernel changes on one box, tools/)
_LIB_PATH = Path(os.environ.get("FLEXS_B200_LIB") or Path(__file__).resolve().parent / "libflexs_b200.so")
_lib: Optional[ctypes.CDLL] = None

OK, EINVAL, ECUDA, EALPHABET = 0, -1, -2, -3
VARIANT_AUTO, VARIANT_SIMPLE, VARIANT_TILED, VARIANT_UMMA, VARIANT_UMMA_LUT, VARIANT_ENUM = 0, 1, 2, 3, 4, 5
VARIANT_NAMES = {0: "auto", 1: "simple", 2: "tiled_ffma", 3: "umma_tcgen05", 4: "lut9_umma_tcgen05", 5: "enum_table"}

#: every symbol include/flexs_b200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("flexs_abi_version", c_int, []),
    ("flexs_last_error", c_char_p, []),
    ("flexs_device_count", c_int, []),
    ("flexs_cnn_create", c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    ("flexs_mlp_create", c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    ("flexs_model_destroy", None, [c_void_p]),
    ("flexs_model_num_arrays", c_int, [c_void_p]),
    ("flexs_model_array_size", c_int64, [c_void_p, c_int]),
    ("flexs_model_set_weights", c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    ("flexs_model_get_weights", c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    ("flexs_model_set_variant", c_int, [c_void_p, c_int]),
    ("flexs_model_active_variant", c_int, [c_void_p, c_int64]),
    ("flexs_model_launch_count", c_int64, [c_void_p]),
    ("flexs_encode_dev", c_int, [c_void_p, c_int64, c_char_p, c_int, c_void_p, c_void_p, c_void_p]),
    ("flexs_model_forward_dev", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    ("flexs_model_score_host", c_int, [c_void_p, c_void_p, c_int64, c_char_p, c_void_p, POINTER(c_int64)]),
    ("flexs_bits_per_residue", c_int, [c_int]),
    ("flexs_packed_row_bytes", c_int64, [c_int, c_int]),
    ("flexs_unpack_dev", c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    ("flexs_pack_dev", c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    ("flexs_model_score_host_packed", c_int, [c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_int64)]),
    ("flexs_topk_workspace_bytes", c_int64, [c_int64, c_int]),
    ("flexs_topk_dev", c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("flexs_topk_select_workspace_bytes", c_int64, []),
    ("flexs_topk_select_dev", c_int, [c_void_p, c_int64, c_int, c_int64, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    ("flexs_screen_message_bytes", c_int64, [c_int, c_int]),
    ("flexs_screen_merge_dev", c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("flexs_peer_mailbox_bytes", c_int64, [c_int64, c_int, c_int]),
    ("flexs_peer_alloc", c_int, [c_int64, ctypes.POINTER(c_void_p), c_char_p]),
    ("flexs_peer_open", c_int, [c_char_p, ctypes.POINTER(c_void_p)]),
    ("flexs_peer_close", c_int, [c_void_p]),
    ("flexs_peer_free", c_int, [c_void_p]),
    ("flexs_screen_push_dev", c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_int, ctypes.c_uint32, c_void_p, c_void_p]),
    ("flexs_screen_wait_dev", c_int, [c_void_p, c_int64, c_int, c_int, c_int, ctypes.c_uint32, c_void_p, c_void_p]),
    ("flexs_dedup_workspace_bytes", c_int64, [c_int64]),
    ("flexs_dedup_scores_dev", c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    ("flexs_dedup_representatives_dev", c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    ("flexs_mutate_dev", c_int, [c_void_p, c_int64, c_int, c_int, c_float, c_uint64, c_uint64, c_void_p, c_void_p]),
    ("flexs_argmax_decode_dev", c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p]),
    ("flexs_edit_density_dev", c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    ("flexs_model_fit_dev", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_uint64, c_void_p, c_void_p]),
    ("flexs_model_train_step_dev", c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, POINTER(c_float), c_void_p]),
    ("flexs_model_get_optimizer_state", c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64)]),
    ("flexs_model_reset_optimizer", c_int, [c_void_p]),
    ("flexs_vae_create", c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    ("flexs_vae_destroy", None, [c_void_p]),
    ("flexs_vae_num_arrays", c_int, [c_void_p]),
    ("flexs_vae_array_size", c_int64, [c_void_p, c_int]),
    ("flexs_vae_set_weights", c_int, [c_void_p, POINTER(c_void_p)]),
    ("flexs_vae_get_weights", c_int, [c_void_p, POINTER(c_void_p)]),
    ("flexs_vae_get_gradients", c_int, [c_void_p, POINTER(c_void_p)]),
    ("flexs_vae_reset_optimizer", c_int, [c_void_p]),
    ("flexs_vae_train_step_dev", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, POINTER(c_float), c_void_p]),
    ("flexs_vae_fit_dev", c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_uint64, c_void_p, POINTER(c_int), c_void_p]),
    ("flexs_vae_decode_dev", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    ("flexs_vae_log_prob_dev", c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    ("flexs_additive_score_dev", c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_double, c_double, c_void_p,
                                         c_void_p, c_void_p]),
    ("flexs_lookup_score_dev", c_int, [c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_int64, c_void_p, c_void_p]),
]

EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]


class NativeError(RuntimeError):
    """A call into libflexs_b200 failed."""


def lib_path() -> Path:
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise NativeError(
                f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `python flexs_b200/_build.py`).  flexs_b200 has no CPU fallback."
            )
        handle = ctypes.CDLL(str(_LIB_PATH), mode=getattr(os, "RTLD_NOW", 2) | getattr(os, "RTLD_LOCAL", 0))
        for name, restype, argtypes in _SIGNATURES:
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    """Turn a negative return code into the Python exception the reference would raise."""
    if rc >= 0:
        return
    msg = lib().flexs_last_error().decode("utf-8", "replace")
    if rc == EALPHABET:
        raise ValueError(msg)  # str.index raises ValueError in the reference (sequence_utils.py:46)
    if rc == EINVAL:
        raise ValueError(f"{what}: {msg}" if what else msg)
    raise NativeError(f"{what}: {msg}" if what else msg)


def device_count() -> int:
    n = lib().flexs_device_count()
    return max(n, 0)


def require_cuda() -> None:
    n = lib().flexs_device_count()
    if n <= 0:
        raise NativeError("flexs_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback")


def _as_ptr_array(arrays: Sequence[np.ndarray]):
    arr_t = c_void_p * len(arrays)
    return arr_t(*[a.ctypes.data_as(c_void_p).value for a in arrays])


class NativeModel:
    """Owns one ``flexs_model_t*`` (a CNN or MLP surrogate with ``n_members`` weight sets)."""

    def __init__(self, kind: str, *, seq_len: int, alphabet_size: int, hidden_size: int, num_filters: int = 0,
                 kernel_size: int = 0, n_members: int = 1, device: int = 0):
        require_cuda()
        self.kind = kind
        self.seq_len, self.alphabet_size, self.hidden_size = seq_len, alphabet_size, hidden_size
        self.num_filters, self.kernel_size, self.n_members, self.device = num_filters, kernel_size, n_members, device
        handle = c_void_p()
        if kind == "cnn":
            rc = lib().flexs_cnn_create(device, seq_len, alphabet_size, num_filters, hidden_size, kernel_size,
                                        n_members, ctypes.byref(handle))
        elif kind == "mlp":
            rc = lib().flexs_mlp_create(device, seq_len, alphabet_size, hidden_size, n_members, ctypes.byref(handle))
        else:
            raise ValueError(kind)
        check(rc, f"create {kind}")
        self._h = handle
        self.array_sizes = [int(lib().flexs_model_array_size(self._h, i))
                            for i in range(lib().flexs_model_num_arrays(self._h))]

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib().flexs_model_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - interpreter shutdown ordering
        try:
            self.close()
        except Exception:
            pass

    # -- weights -------------------------------------------------------------------------
    def set_weights(self, weights: Sequence[np.ndarray], member: int = 0) -> None:
        if len(weights) != len(self.array_sizes):
            raise ValueError(f"expected {len(self.array_sizes)} arrays, got {len(weights)}")
        flat = []
        for w, size in zip(weights, self.array_sizes):
            a = np.ascontiguousarray(np.asarray(w, dtype=np.float32)).reshape(-1)
            if a.size != size:
                raise ValueError(f"weight array has {a.size} elements, expected {size}")
            flat.append(a)
        check(lib().flexs_model_set_weights(self._h, member, _as_ptr_array(flat)), "set_weights")

    def get_weights(self, member: int = 0) -> List[np.ndarray]:
        flat = [np.empty(size, dtype=np.float32) for size in self.array_sizes]
        check(lib().flexs_model_get_weights(self._h, member, _as_ptr_array(flat)), "get_weights")
        return flat

    # -- variants / counters ---------------------------------------------------------------
    def set_variant(self, variant: int) -> None:
        check(lib().flexs_model_set_variant(self._h, variant), "set_variant")

    def active_variant(self, n: int = 1) -> int:
        return int(lib().flexs_model_active_variant(self._h, n))

    @property
    def launch_count(self) -> int:
        return int(lib().flexs_model_launch_count(self._h))

    # -- forward ---------------------------------------------------------------------------
    def forward_dev(self, d_idx: int, n: int, d_out: int, stream: int = 0) -> None:
        """Enqueue scoring of ``uint8[n, L]`` at device address ``d_idx`` into ``float32[n]`` at ``d_out``."""
        check(lib().flexs_model_forward_dev(self._h, c_void_p(d_idx), n, c_void_p(d_out), c_void_p(stream)), "forward")

    def score_host(self, chars: np.ndarray, alphabet: str, out: Optional[np.ndarray] = None) -> np.ndarray:
        """``chars``: contiguous ``uint8[n, L]`` of residue CHARACTERS in host memory -> ``float32[n]``."""
        if chars.dtype != np.uint8 or not chars.flags.c_contiguous:
            raise ValueError("chars must be a C-contiguous uint8 array")
        n = chars.shape[0] if chars.ndim == 2 else chars.size // max(self.seq_len, 1)
        if chars.size != n * self.seq_len:
            raise ValueError(f"expected sequences of length {self.seq_len}")
        if out is None:
            out = np.empty(n, dtype=np.float32)
        bad = c_int64(-1)
        alpha = alphabet if isinstance(alphabet, bytes) else alphabet.encode("latin-1")
        rc = lib().flexs_model_score_host(self._h, chars.ctypes.data_as(c_void_p), n, alpha,
                                          out.ctypes.data_as(c_void_p), ctypes.byref(bad))
        if rc == EALPHABET:
            pos = bad.value
            raise ValueError(f"substring not found: character {chr(int(chars.reshape(-1)[pos]))!r} of sequence "
                             f"{pos // self.seq_len} (position {pos % self.seq_len}) is not in alphabet {alphabet!r}")
        check(rc, "score_host")
        return out

    def score_host_packed(self, packed: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """``packed``: contiguous ``uint8[n, ceil(L * bits / 8)]`` rows in the wire format of include/flexs_b200.h
        (``sequence_utils.pack_sequences``) in host memory -> ``float32[n]``."""
        row_bytes = packed_row_bytes(self.seq_len, self.alphabet_size)
        if packed.dtype != np.uint8 or not packed.flags.c_contiguous or packed.ndim != 2 or packed.shape[1] != row_bytes:
            raise ValueError(f"packed must be a C-contiguous uint8 [n, {row_bytes}] array")
        n = packed.shape[0]
        if out is None:
            out = np.empty(n, dtype=np.float32)
        bad = c_int64(-1)
        rc = lib().flexs_model_score_host_packed(self._h, packed.ctypes.data_as(c_void_p), n,
                                                 out.ctypes.data_as(c_void_p), ctypes.byref(bad))
        if rc == EALPHABET:
            pos = bad.value
            raise ValueError(f"substring not found: residue {pos % self.seq_len} of sequence {pos // self.seq_len} "
                             f"is not a valid index into the alphabet")
        check(rc, "score_host_packed")
        return out

    # -- training --------------------------------------------------------------------------
    def fit_dev(self, d_idx: int, d_labels: int, n: int, batch_size: int, epochs: int, seed: int,
                stream: int = 0) -> np.ndarray:
        losses = np.zeros(max(epochs, 1) * self.n_members, dtype=np.float32)
        check(lib().flexs_model_fit_dev(self._h, c_void_p(d_idx), c_void_p(d_labels), n, batch_size, epochs,
                                        c_uint64(seed), losses.ctypes.data_as(c_void_p), c_void_p(stream)), "fit")
        return losses

    def optimizer_state(self, member: int = 0):
        """(m arrays, v arrays, step) of the Adam optimiser of one member."""
        ms = [np.empty(size, dtype=np.float32) for size in self.array_sizes]
        vs = [np.empty(size, dtype=np.float32) for size in self.array_sizes]
        step = c_int64(0)
        check(lib().flexs_model_get_optimizer_state(self._h, member, _as_ptr_array(ms), _as_ptr_array(vs),
                                                    ctypes.byref(step)), "get_optimizer_state")
        return ms, vs, int(step.value)

    def reset_optimizer(self) -> None:
        check(lib().flexs_model_reset_optimizer(self._h), "reset_optimizer")

    def train_step_dev(self, member: int, d_idx: int, d_labels: int, n: int, d_mask: int = 0, stream: int = 0) -> float:
        loss = c_float(0.0)
        check(lib().flexs_model_train_step_dev(self._h, member, c_void_p(d_idx), c_void_p(d_labels), n,
                                               c_void_p(d_mask), ctypes.byref(loss), c_void_p(stream)), "train_step")
        return float(loss.value)


class NativeVAE:
    """Owns one ``flexs_vae_t*`` (K9: the CbAS / DbAS generator; see include/flexs_b200.h)."""

    def __init__(self, seq_len: int, alphabet_size: int, intermediate_dim: int, latent_dim: int, device: int = 0):
        require_cuda()
        self.seq_len, self.alphabet_size, self.intermediate_dim, self.latent_dim, self.device = \
            seq_len, alphabet_size, intermediate_dim, latent_dim, device
        handle = c_void_p()
        check(lib().flexs_vae_create(device, seq_len, alphabet_size, intermediate_dim, latent_dim, ctypes.byref(handle)), "vae_create")
        self._h = handle
        self.array_sizes = [int(lib().flexs_vae_array_size(self._h, i)) for i in range(lib().flexs_vae_num_arrays(self._h))]
        d, i, z = seq_len * alphabet_size, intermediate_dim, latent_dim
        self.array_shapes = [(d, i), (i,), (i, i), (i,), (i,), (i,), (i,), (i,), (i, i), (i,), (i, z), (z,), (i, z), (z,),
                             (z, i), (i,), (i, i), (i,), (i, i), (i,), (i, d), (d,)]

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib().flexs_vae_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, weights: Sequence[np.ndarray]) -> None:
        if len(weights) != len(self.array_sizes):
            raise ValueError(f"expected {len(self.array_sizes)} arrays, got {len(weights)}")
        flat = []
        for w, size in zip(weights, self.array_sizes):
            a = np.ascontiguousarray(np.asarray(w, dtype=np.float32)).reshape(-1)
            if a.size != size:
                raise ValueError(f"weight array has {a.size} elements, expected {size}")
            flat.append(a)
        check(lib().flexs_vae_set_weights(self._h, _as_ptr_array(flat)), "vae_set_weights")

    def _read(self, fn, what) -> List[np.ndarray]:
        flat = [np.empty(size, dtype=np.float32) for size in self.array_sizes]
        check(fn(self._h, _as_ptr_array(flat)), what)
        return [a.reshape(shp) for a, shp in zip(flat, self.array_shapes)]

    def get_weights(self) -> List[np.ndarray]:
        return self._read(lib().flexs_vae_get_weights, "vae_get_weights")

    def get_gradients(self) -> List[np.ndarray]:
        return self._read(lib().flexs_vae_get_gradients, "vae_get_gradients")

    def reset_optimizer(self) -> None:
        check(lib().flexs_vae_reset_optimizer(self._h), "vae_reset_optimizer")

    def train_step_dev(self, d_idx: int, d_weights: int, n: int, d_mask1: int, d_mask2: int, d_eps: int, stream: int = 0) -> float:
        loss = c_float(0.0)
        check(lib().flexs_vae_train_step_dev(self._h, c_void_p(d_idx), c_void_p(d_weights), n, c_void_p(d_mask1), c_void_p(d_mask2),
                                             c_void_p(d_eps), ctypes.byref(loss), c_void_p(stream)), "vae_train_step")
        return float(loss.value)

    def fit_dev(self, d_idx: int, d_weights: int, n_train: int, batch_size: int, epochs: int, patience: int, seed: int,
                stream: int = 0):
        losses = np.zeros(max(epochs, 1), dtype=np.float32)
        ran = c_int(0)
        check(lib().flexs_vae_fit_dev(self._h, c_void_p(d_idx), c_void_p(d_weights), n_train, batch_size, epochs, patience,
                                      c_uint64(seed & 0xFFFFFFFFFFFFFFFF), losses.ctypes.data_as(c_void_p), ctypes.byref(ran),
                                      c_void_p(stream)), "vae_fit")
        return losses[: ran.value], int(ran.value)

    def decode_dev(self, d_z: int, n: int, d_out: int, stream: int = 0) -> None:
        check(lib().flexs_vae_decode_dev(self._h, c_void_p(d_z), n, c_void_p(d_out), c_void_p(stream)), "vae_decode")

    def log_prob_dev(self, d_idx: int, n: int, d_eps: int, d_logp: int, stream: int = 0) -> None:
        check(lib().flexs_vae_log_prob_dev(self._h, c_void_p(d_idx), n, c_void_p(d_eps), c_void_p(d_logp), c_void_p(stream)),
              "vae_log_prob")


# -- stateless kernels ------------------------------------------------------------------------
def encode_dev(d_chars: int, n_bytes: int, alphabet: str, d_idx: int, d_status: int, stream: int = 0) -> None:
    check(lib().flexs_encode_dev(c_void_p(d_chars), n_bytes, alphabet.encode("latin-1"), len(alphabet),
                                 c_void_p(d_idx), c_void_p(d_status), c_void_p(stream)), "encode")


def bits_per_residue(alphabet_size: int) -> int:
    bits = 1
    while (1 << bits) < alphabet_size:
        bits += 1
    return bits


def packed_row_bytes(seq_len: int, alphabet_size: int) -> int:
    """Bytes of one sequence in the packed wire format (pure arithmetic, mirrors flexs_packed_row_bytes)."""
    return (seq_len * bits_per_residue(alphabet_size) + 7) // 8


def unpack_dev(d_packed: int, n: int, seq_len: int, alphabet_size: int, d_idx: int, d_status: int, stream: int = 0) -> None:
    check(lib().flexs_unpack_dev(c_void_p(d_packed), n, seq_len, alphabet_size, c_void_p(d_idx), c_void_p(d_status),
                                 c_void_p(stream)), "unpack")


def pack_dev(d_idx: int, n: int, seq_len: int, alphabet_size: int, d_packed: int, stream: int = 0) -> None:
    check(lib().flexs_pack_dev(c_void_p(d_idx), n, seq_len, alphabet_size, c_void_p(d_packed), c_void_p(stream)), "pack")


def topk_workspace_bytes(n: int, k: int) -> int:
    b = int(lib().flexs_topk_workspace_bytes(n, k))
    if b < 0:
        raise ValueError("k must be in [1, 4096]")
    return b


def topk_dev(d_scores: int, n: int, k: int, index_offset: int, d_index_map: int, d_top_scores: int, d_top_idx: int,
             d_work: int, stream: int = 0) -> None:
    check(lib().flexs_topk_dev(c_void_p(d_scores), n, k, index_offset, c_void_p(d_index_map), c_void_p(d_top_scores),
                               c_void_p(d_top_idx), c_void_p(d_work), c_void_p(stream)), "topk")


def topk_select_workspace_bytes() -> int:
    return int(lib().flexs_topk_select_workspace_bytes())


def topk_select_dev(d_scores: int, n: int, k: int, index_offset: int, d_rows: int, row_len: int, unique: bool,
                    d_top_scores: int, d_top_idx: int, d_top_rows: int, d_status: int, d_work: int, stream: int = 0) -> None:
    """Single-launch exact top-k (optionally over distinct rows, with the winners' rows); see include/flexs_b200.h."""
    check(lib().flexs_topk_select_dev(c_void_p(d_scores), n, k, index_offset, c_void_p(d_rows), row_len, 1 if unique else 0,
                                      c_void_p(d_top_scores), c_void_p(d_top_idx), c_void_p(d_top_rows), c_void_p(d_status),
                                      c_void_p(d_work), c_void_p(stream)), "topk_select")


def screen_message_bytes(k: int, seq_len: int) -> int:
    return int(lib().flexs_screen_message_bytes(k, seq_len))


def screen_merge_dev(d_gathered: int, world: int, k: int, seq_len: int, d_top_scores: int, d_top_idx: int,
                     d_top_rows: int, stream: int = 0) -> None:
    check(lib().flexs_screen_merge_dev(c_void_p(d_gathered), world, k, seq_len, c_void_p(d_top_scores), c_void_p(d_top_idx),
                                       c_void_p(d_top_rows), c_void_p(stream)), "screen_merge")


def peer_mailbox_bytes(msg_bytes: int, world: int, depth: int) -> int:
    b = int(lib().flexs_peer_mailbox_bytes(msg_bytes, world, depth))
    if b < 0:
        raise ValueError("msg_bytes must be a positive multiple of 16, world and depth >= 1")
    return b


def peer_alloc(nbytes: int):
    """cudaMalloc'ed, zeroed device buffer + its 64-byte cudaIpc handle: ``(device pointer, handle bytes)``."""
    ptr, handle = c_void_p(), ctypes.create_string_buffer(64)
    check(lib().flexs_peer_alloc(nbytes, ctypes.byref(ptr), handle), "peer_alloc")
    return int(ptr.value), handle.raw


def peer_open(handle: bytes) -> int:
    """Map another process's buffer (same node) into this process; returns the device pointer."""
    ptr = c_void_p()
    check(lib().flexs_peer_open(ctypes.create_string_buffer(handle, 64), ctypes.byref(ptr)), "peer_open")
    return int(ptr.value)


def peer_close(ptr: int) -> None:
    check(lib().flexs_peer_close(c_void_p(ptr)), "peer_close")


def peer_free(ptr: int) -> None:
    check(lib().flexs_peer_free(c_void_p(ptr)), "peer_free")


def screen_push_dev(d_msg: int, msg_bytes: int, rank: int, world: int, slot: int, depth: int, seq: int, d_peer_bases: int,
                    stream: int = 0) -> None:
    check(lib().flexs_screen_push_dev(c_void_p(d_msg), msg_bytes, rank, world, slot, depth, seq, c_void_p(d_peer_bases),
                                      c_void_p(stream)), "screen_push")


def screen_wait_dev(d_mailbox: int, msg_bytes: int, world: int, slot: int, depth: int, seq: int, d_status: int = 0,
                    stream: int = 0) -> None:
    check(lib().flexs_screen_wait_dev(c_void_p(d_mailbox), msg_bytes, world, slot, depth, seq, c_void_p(d_status),
                                      c_void_p(stream)), "screen_wait")


def dedup_workspace_bytes(n: int) -> int:
    b = int(lib().flexs_dedup_workspace_bytes(n))
    if b < 0:
        raise ValueError("n must be in [0, 2^31)")
    return b


def dedup_scores_dev(d_idx: int, n: int, seq_len: int, d_scores: int, d_scores_out: int, d_work: int, stream: int = 0) -> None:
    """scores_out[i] = scores[i] if row i is the first occurrence of its sequence, else -inf (exact)."""
    check(lib().flexs_dedup_scores_dev(c_void_p(d_idx), n, seq_len, c_void_p(d_scores), c_void_p(d_scores_out),
                                       c_void_p(d_work), c_void_p(stream)), "dedup")


def dedup_representatives_dev(d_idx: int, n: int, seq_len: int, d_rep: int, d_work: int, stream: int = 0) -> None:
    """rep[i] = lowest index of a row equal to row i (exact); workspace as for ``dedup_scores_dev``."""
    check(lib().flexs_dedup_representatives_dev(c_void_p(d_idx), n, seq_len, c_void_p(d_rep), c_void_p(d_work),
                                                c_void_p(stream)), "dedup_representatives")


def mutate_dev(d_parents: int, n: int, seq_len: int, alphabet_size: int, mu: float, seed: int, subsequence: int,
               d_children: int, stream: int = 0) -> None:
    check(lib().flexs_mutate_dev(c_void_p(d_parents), n, seq_len, alphabet_size, c_float(mu), c_uint64(seed),
                                 c_uint64(subsequence), c_void_p(d_children), c_void_p(stream)), "mutate")


def argmax_decode_dev(d_x: int, n: int, seq_len: int, row_stride: int, alphabet_size: int, d_idx: int,
                      stream: int = 0) -> None:
    check(lib().flexs_argmax_decode_dev(c_void_p(d_x), n, seq_len, row_stride, alphabet_size, c_void_p(d_idx),
                                        c_void_p(stream)), "argmax_decode")


def edit_density_dev(d_new: int, n_new: int, d_seen: int, d_seen_fitness: int, n_seen: int, seq_len: int, radius: int,
                     d_out: int, stream: int = 0) -> None:
    """K8: density penalty of the DyNA-PPO environment for a batch (environments/dyna_ppo.py:106-114)."""
    check(lib().flexs_edit_density_dev(c_void_p(d_new), n_new, c_void_p(d_seen), c_void_p(d_seen_fitness), n_seen, seq_len,
                                       radius, c_void_p(d_out), c_void_p(stream)), "edit_density")


def _column_table(column_of_char: Optional[np.ndarray]):
    if column_of_char is None:
        return None, c_void_p(0)
    lut = np.ascontiguousarray(column_of_char, dtype=np.uint8)
    if lut.shape != (256,):
        raise ValueError("column_of_char must be uint8[256]")
    return lut, lut.ctypes.data_as(c_void_p)


def additive_score_dev(d_seq: int, n: int, seq_len: int, column_of_char: Optional[np.ndarray], ncols: int, d_table: int,
                       offset: float, denom: float, d_noise: int, d_out: int, stream: int = 0) -> None:
    """K6 additive table landscape (additive_aav_packaging.py:101-118) on device buffers."""
    keep, ptr = _column_table(column_of_char)
    check(lib().flexs_additive_score_dev(c_void_p(d_seq), n, seq_len, ptr, ncols, c_void_p(d_table), offset, denom,
                                         c_void_p(d_noise), c_void_p(d_out), c_void_p(stream)), "additive_score")


def lookup_score_dev(d_seq: int, n: int, seq_len: int, column_of_char: Optional[np.ndarray], base: int, d_table: int,
                     table_len: int, d_out: int, stream: int = 0) -> None:
    """K6 dictionary landscape (tf_binding.py:43-44) as a dense table on device buffers."""
    keep, ptr = _column_table(column_of_char)
    check(lib().flexs_lookup_score_dev(c_void_p(d_seq), n, seq_len, ptr, base, c_void_p(d_table), table_len,
                                       c_void_p(d_out), c_void_p(stream)), "lookup_score")
