"""VAE generator for CbAS/DbAS (reference: flexs/utils/VAE_utils.py:12-232).

§8(f) "next row" #1: the generator is not the surrogate hot path, but CbAS cannot run without it, it dominates the
explorer's wall-clock (a refit after every 100 proposals), and the reference's is a Keras model.  Same class,
constructor arguments and methods (``train_model``, ``generate``, ``calculate_log_probability``) and the same
architecture / losses / sampling procedure.  With a CUDA device the network is K9 of libflexs_b200
(csrc/vae.cu: ``flexs_vae_*`` — the whole fit runs on the device with one host synchronisation per epoch, the
decoder pass and the reconstruction log-probability are single calls on residue indices); without one the torch
module below stands in so the explorer logic stays testable on a CPU box (it is not a product path).

Reference behaviours kept on purpose:
  * encoder Dense(elu) -> Dropout(0.3) -> Dense(elu) -> BatchNorm -> Dense(elu) -> (z_mean, z_log_var);
    decoder Dense(elu) x2 -> Dropout(0.3) -> Dense(elu) -> Dense(sigmoid)            (VAE_utils.py:40-63)
  * loss = original_dim * mean(BCE) + KL, Adam(lr 1e-4, clipvalue 0.5), last ``validation_split`` of the data
    held out, early stopping on the training loss with patience 3                      (:75-92, :124-151)
  * ``generate`` decodes ONE latent sample, reshapes the flat (L*A) output to (A, L) *without transposing*
    (:158-160 — a quirk of the reference), converts to Boltzmann weights at temperature 1e-3 and rejection-samples
    new sequences, multiplying the temperature by 1.3 on every repeat                   (:153-187)
  * ``calculate_log_probability``: per-residue reconstruction probability, sum of logs, nan_to_num (:189-217)
"""
import random
from typing import List, Optional

import numpy as np
import scipy.special
import torch
from torch import nn

from flexs_b200.types import SEQUENCES_TYPE
from flexs_b200.utils import sequence_utils as s_utils


def _device() -> torch.device:
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


class VAEModel(nn.Module):
    """Encoder/decoder pair with the reference's layer stack."""

    def __init__(self, original_dim: int, intermediate_dim: int, latent_dim: int):
        super().__init__()
        self.original_dim, self.latent_dim = original_dim, latent_dim
        h = intermediate_dim
        self.enc = nn.Sequential(nn.Linear(original_dim, h), nn.ELU(), nn.Dropout(0.3), nn.Linear(h, h), nn.ELU(),
                                 nn.BatchNorm1d(h, eps=1e-3, momentum=0.01), nn.Linear(h, h), nn.ELU())
        self.z_mean = nn.Linear(h, latent_dim)
        self.z_log_var = nn.Linear(h, latent_dim)
        self.dec = nn.Sequential(nn.Linear(latent_dim, h), nn.ELU(), nn.Linear(h, h), nn.ELU(), nn.Dropout(0.3),
                                 nn.Linear(h, h), nn.ELU(), nn.Linear(h, original_dim), nn.Sigmoid())
        for mod in self.modules():  # Keras defaults: glorot-uniform kernels, zero biases
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                nn.init.zeros_(mod.bias)

    def encode(self, x):
        hdn = self.enc(x)
        return self.z_mean(hdn), self.z_log_var(hdn)

    def forward(self, x):
        mean, log_var = self.encode(x)
        z = mean + torch.exp(0.5 * log_var) * torch.randn_like(mean)
        return self.dec(z)

    def predict(self, x: np.ndarray) -> np.ndarray:
        """Keras ``predict``: inference-mode reconstruction (a latent sample is still drawn, :65-69)."""
        self.eval()
        with torch.no_grad():
            dev = next(self.parameters()).device
            return self(torch.as_tensor(x, dtype=torch.float32, device=dev)).cpu().numpy()

    def generate(self) -> np.ndarray:
        """Decode one standard-normal latent sample (:71-74)."""
        self.eval()
        with torch.no_grad():
            dev = next(self.parameters()).device
            z = torch.as_tensor(np.random.randn(1, self.latent_dim), dtype=torch.float32, device=dev)
            return self.dec(z).cpu().numpy()

    # keras-style weight cloning used by CbAS (cbas_dbas.py:130-144)
    def get_weights(self) -> List[np.ndarray]:
        return [v.detach().cpu().numpy().copy() for v in self.state_dict().values()]

    def set_weights(self, weights: List[np.ndarray]) -> None:
        state = self.state_dict()
        for (name, old), new in zip(state.items(), weights):
            state[name] = torch.as_tensor(new, dtype=old.dtype, device=old.device).reshape(old.shape)
        self.load_state_dict(state)


class NativeVAEModel:
    """``VAEModel`` on the B200 kernels (K9): owns a ``flexs_vae_t`` and mirrors the Keras model's surface that
    CbAS touches — ``get_weights`` / ``set_weights`` (cbas_dbas.py:130-144 clones the prior), ``generate`` and ``predict``."""

    def __init__(self, seq_length: int, alphabet_size: int, intermediate_dim: int, latent_dim: int, device: int = 0,
                 seed: Optional[int] = None):
        from flexs_b200 import _native

        self.seq_length, self.alphabet_size = seq_length, alphabet_size
        self.original_dim, self.latent_dim, self.device = seq_length * alphabet_size, latent_dim, device
        self.native = _native.NativeVAE(seq_length, alphabet_size, intermediate_dim, latent_dim, device)
        rng = np.random.default_rng(seed)
        weights = []
        for i, shp in enumerate(self.native.array_shapes):   # Keras defaults: glorot-uniform kernels, zero biases, BN (1, 0, 0, 1)
            if len(shp) == 2:
                lim = np.sqrt(6.0 / (shp[0] + shp[1]))
                weights.append(rng.uniform(-lim, lim, size=shp).astype(np.float32))
            else:
                weights.append(np.ones(shp, np.float32) if i in (4, 7) else np.zeros(shp, np.float32))
        self.native.set_weights(weights)

    def get_weights(self) -> List[np.ndarray]:
        return self.native.get_weights()

    def set_weights(self, weights: List[np.ndarray]) -> None:
        self.native.set_weights(weights)

    def _dev(self):
        return torch.device("cuda", self.device)

    def generate(self) -> np.ndarray:
        """Decode one standard-normal latent sample (:71-74)."""
        z = torch.as_tensor(np.random.randn(1, self.latent_dim), dtype=torch.float32, device=self._dev())
        out = torch.empty((1, self.original_dim), dtype=torch.float32, device=self._dev())
        with torch.cuda.device(self._dev()):
            self.native.decode_dev(z.data_ptr(), 1, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out.cpu().numpy()

    def log_probability(self, idx: np.ndarray) -> np.ndarray:
        """calculate_log_probability (:189-217) on residue indices; the latent noise ``predict`` draws comes from numpy."""
        n = len(idx)
        dev = self._dev()
        d_idx = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.uint8)).to(dev)
        eps = torch.as_tensor(np.random.randn(n, self.latent_dim), dtype=torch.float32, device=dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            self.native.log_prob_dev(d_idx.data_ptr(), n, eps.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out.cpu().numpy()


class VAE:
    """Wrapper exposing the interface CbAS/DbAS use."""

    def __init__(self, seq_length: int, alphabet: str, batch_size: int = 10, latent_dim: int = 2,
                 intermediate_dim: int = 250, epochs: int = 10, epsilon_std: float = 1.0, beta: float = 1,
                 validation_split: float = 0.2, verbose: bool = True, device: Optional[int] = None,
                 seed: Optional[int] = None):
        self.batch_size, self.latent_dim, self.intermediate_dim = batch_size, latent_dim, intermediate_dim
        self.epochs, self.epsilon_std, self.beta = epochs, epsilon_std, beta
        self.validation_split, self.verbose = validation_split, verbose
        self.name = f"VAE_latent_dim={latent_dim}_intermediate_dim={intermediate_dim}"
        self.alphabet, self.seq_length = alphabet, seq_length
        self.device, self.seed = device, seed
        self._fits = 0
        self.last_fit_losses = None
        if torch.cuda.is_available():
            self.vae = NativeVAEModel(seq_length, len(alphabet), intermediate_dim, latent_dim,
                                      torch.cuda.current_device() if device is None else device, seed)
            self._opt = None
        else:   # GPU-less box: torch stand-in, explorer logic only
            self.vae = VAEModel(len(alphabet) * seq_length, intermediate_dim, latent_dim).to(_device())
            self._opt = torch.optim.Adam(self.vae.parameters(), lr=1e-4, eps=1e-7)

    @property
    def native(self) -> bool:
        return isinstance(self.vae, NativeVAEModel)

    def _train_native(self, samples, weights):
        """``fit`` (:141-151) on the device: the LAST ``validation_split`` of the data is held out (Keras takes it before
        shuffling), mini-batches of ``batch_size`` reshuffled every epoch, early stopping on the training loss, patience 3."""
        idx = s_utils.encode_sequences(list(samples), self.alphabet)
        n_train = len(idx) - int(len(idx) * self.validation_split)
        dev = torch.device("cuda", self.vae.device)
        d_idx = torch.from_numpy(np.ascontiguousarray(idx[:n_train])).to(dev)
        d_w = torch.as_tensor(np.asarray(weights, dtype=np.float32)[:n_train], device=dev)
        seed = (0 if self.seed is None else int(self.seed)) * 1000003 + self._fits + 1
        self._fits += 1
        with torch.cuda.device(dev):
            losses, ran = self.vae.native.fit_dev(d_idx.data_ptr(), d_w.data_ptr(), n_train, max(2, self.batch_size), self.epochs, 3,
                                                  seed, torch.cuda.current_stream().cuda_stream)
        self.last_fit_losses = losses
        if self.verbose:
            for e, l in enumerate(losses):
                print(f"Epoch {e + 1}/{self.epochs} - loss: {l:.4f}")

    def _one_hots(self, sequences) -> np.ndarray:
        idx = s_utils.encode_sequences(list(sequences), self.alphabet)
        out = np.zeros((len(idx), self.seq_length, len(self.alphabet)), dtype=np.float32)
        np.put_along_axis(out, idx[..., None].astype(np.int64), 1.0, axis=2)
        return out

    def train_model(self, samples, weights):
        """Fit on ``samples`` weighted by ``weights`` (:132-151)."""
        if self.native:
            return self._train_native(samples, weights)
        x = self._one_hots(samples).reshape(len(samples), -1)
        w = np.asarray(weights, dtype=np.float32)
        n_train = len(x) - int(len(x) * self.validation_split)   # Keras holds out the LAST fraction
        dev = next(self.vae.parameters()).device
        x_t = torch.as_tensor(x[:n_train], device=dev)
        w_t = torch.as_tensor(w[:n_train], device=dev)
        best, bad = float("inf"), 0
        for epoch in range(self.epochs):
            self.vae.train()
            perm = torch.randperm(n_train, device=dev)
            total, count = 0.0, 0
            for start in range(0, n_train, self.batch_size):
                sel = perm[start: start + self.batch_size]
                if len(sel) < 2:  # BatchNorm needs more than one sample in training mode
                    continue
                xb, wb = x_t[sel], w_t[sel]
                mean, log_var = self.vae.encode(xb)
                z = mean + torch.exp(0.5 * log_var) * torch.randn_like(mean)
                recon = self.vae.dec(z)
                bce = nn.functional.binary_cross_entropy(recon, xb, reduction="none").mean(dim=1)
                recon_loss = self.vae.original_dim * bce
                kl = -0.5 * (1 + log_var - mean.pow(2) - log_var.exp()).mean(dim=1)
                loss = ((recon_loss + kl) * wb).sum() / max(float(len(sel)), 1.0)
                self._opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_value_(self.vae.parameters(), 0.5)
                self._opt.step()
                total += float(loss.detach()) * len(sel)
                count += len(sel)
            epoch_loss = total / max(count, 1)
            if self.verbose:
                print(f"Epoch {epoch + 1}/{self.epochs} - loss: {epoch_loss:.4f}")
            if epoch_loss < best - 1e-12:
                best, bad = epoch_loss, 0
            else:
                bad += 1
                if bad >= 3:  # EarlyStopping(monitor="loss", patience=3)
                    break

    def generate(self, n_samples, existing_samples, existing_weights) -> List[str]:
        """``n_samples`` new sequences none of which is in ``existing_samples`` (:153-187)."""
        pwm = np.reshape(self.vae.generate(), (len(self.alphabet), self.seq_length))  # reference quirk: no transpose
        if np.isnan(pwm).any() or np.isinf(pwm).any():
            raise ValueError("NaN and/or inf in the reconstruction matrix")
        existing = set(existing_samples)
        proposals, seen = [], set()
        temperature = 0.001
        weights = pwm_to_boltzmann_weights(pwm, temperature)
        while len(proposals) < n_samples:
            new_seq = "".join(random.choices(self.alphabet, weights[:, pos])[0] for pos in range(self.seq_length))
            if new_seq not in seen and new_seq not in existing:
                proposals.append(new_seq)
                seen.add(new_seq)
            else:
                temperature = 1.3 * temperature
                weights = pwm_to_boltzmann_weights(pwm, temperature)
        return proposals

    def calculate_log_probability(self, sequences: SEQUENCES_TYPE, vae: Optional[VAEModel] = None):
        """log-probability of reconstructing each sequence (:189-217)."""
        vae = vae or self.vae
        if isinstance(vae, NativeVAEModel):
            return vae.log_probability(s_utils.encode_sequences(list(sequences), self.alphabet))
        one_hots = self._one_hots(sequences)
        decoded = vae.predict(one_hots.reshape(len(one_hots), -1)).reshape(one_hots.shape)
        per_res = (decoded * one_hots).max(axis=2) / decoded.sum(axis=2)
        return np.nan_to_num(np.log(1e-9 + per_res).sum(axis=1))


def pwm_to_boltzmann_weights(prob_weight_matrix, temp):
    """Column-wise softmax of ``pwm / temp`` (:220-232)."""
    w = np.asarray(prob_weight_matrix, dtype=np.float64)
    return np.exp(w / temp - scipy.special.logsumexp(w / temp, axis=0, keepdims=True))
