"""A deterministic stand-in for the VAE generator, shared by tests/golden/make_golden_cbas.py (which runs the REFERENCE's
CbAS.propose_sequences with it) and tests/test_explorers_cpu.py (which runs this repo's).  Everything the explorer can observe —
proposals, log-probabilities, what it was trained on — is a pure function of the call history, so the two explorers see the
same generator and must make the same decisions (gamma, importance weights, masking, sample pool, ranking, cost)."""
import hashlib

import numpy as np


def _h(*parts) -> int:
    m = hashlib.sha256("|".join(str(p) for p in parts).encode()).digest()
    return int.from_bytes(m[:8], "little")


class _State:
    def __init__(self):
        self.state = 0

    def get_weights(self):
        return [np.array([self.state], dtype=np.int64)]

    def set_weights(self, weights):
        self.state = int(np.asarray(weights[0]).reshape(-1)[0])


class FakeVAE:
    def __init__(self, seq_length, alphabet, batch_size=10, latent_dim=2, intermediate_dim=250, epochs=10, epsilon_std=1.0,
                 beta=1, validation_split=0.2, verbose=True, **_):
        self.seq_length, self.alphabet = seq_length, alphabet
        self.batch_size, self.latent_dim, self.intermediate_dim, self.epochs = batch_size, latent_dim, intermediate_dim, epochs
        self.epsilon_std, self.beta, self.validation_split, self.verbose = epsilon_std, beta, validation_split, verbose
        self.name = f"VAE_latent_dim={latent_dim}_intermediate_dim={intermediate_dim}"
        self.vae = _State()
        self.train_log = []      # (number of samples, rounded weight sum, number of zero weights, digest of the samples)

    def train_model(self, samples, weights):
        samples, weights = list(samples), np.asarray(weights, dtype=np.float64)
        digest = _h(*samples) % (1 << 32)
        self.train_log.append([len(samples), round(float(weights.sum()), 9), int((weights == 0).sum()), digest])
        self.vae.state = _h(self.vae.state, len(samples), round(float(weights.sum()), 9), digest) % (1 << 40)

    def generate(self, n_samples, existing_samples, existing_weights):
        existing, out, k = set(existing_samples), [], 0
        while len(out) < n_samples:
            rng = np.random.default_rng(_h(self.vae.state, "gen", k))
            seq = "".join(self.alphabet[i] for i in rng.integers(0, len(self.alphabet), size=self.seq_length))
            k += 1
            if seq not in existing and seq not in out:
                out.append(seq)
        return out

    def calculate_log_probability(self, sequences, vae=None):
        state = (vae or self.vae).state
        return np.array([-(_h(state, "lp", s) % 100000) / 10000.0 for s in sequences])
