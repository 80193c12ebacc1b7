"""DyNA-PPO explorer (reference: flexs/baselines/explorers/dyna_ppo.py:32-319 and
environments/dyna_ppo.py:13-163).

§8(f) "next row" #2.  The reference builds on tf-agents (PPOAgent, DynamicEpisodeDriver, replay buffer), which
is absent; this is a restatement of the same loop with a small torch PPO (actor and value MLPs with one
128-unit layer, Adam 1e-5, 10 epochs per update — dyna_ppo.py:208-232).  What is kept exactly is everything
the surrogate hot path sees and everything that determines the bookkeeping:

  * constructive episodes: one residue per step for ``env_batch_size`` parallel sequences; the episode ends
    after L-1 steps, so the LAST position is never sampled and decodes to ``alphabet[0]`` (environments/
    dyna_ppo.py:137, :144-147 — reference quirk);
  * at the end of an episode the complete sequences are decoded by per-position argmax over the first A
    channels (the mask channel is dropped) and scored with ONE ``get_fitness`` call — the landscape in the
    experiment-based round, the model in the model-based rounds (:148-152); every call is charged to ``cost``;
  * reward = fitness - 0.1 * density, density = sum over previously seen sequences within edit distance 2 of
    their fitness / distance (:106-114, :155-160);
  * budgets: experiment round until ``landscape.cost`` grew by ``sequences_batch_size``; each of the
    ``num_model_rounds`` model rounds until ``model.cost`` grew by ``model_queries_per_batch / num_model_rounds``
    (dyna_ppo.py:284-307); proposals = the B best unmeasured sequences generated in the model rounds,
    ``np.argsort(preds)[::-1][:B]`` (B items, unlike the other explorers; :309-319).
"""
from typing import Dict, List, Optional, Tuple

import numpy as np
import pandas as pd
import torch
from torch import nn

from flexs_b200.explorer import Explorer
from flexs_b200.landscape import Landscape
from flexs_b200.model import Model
from flexs_b200.utils import sequence_utils as s_utils


def bounded_edit_distance(a: str, b: str, radius: int) -> int:
    """Levenshtein distance, or ``radius + 1`` when it exceeds ``radius`` (banded DP)."""
    if abs(len(a) - len(b)) > radius:
        return radius + 1
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i] + [radius + 1] * len(b)
        lo, hi = max(1, i - radius), min(len(b), i + radius)
        for j in range(lo, hi + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != b[j - 1]))
        if min(cur[max(0, lo - 1): hi + 1]) > radius:
            return radius + 1
        prev = cur
    return min(prev[len(b)], radius + 1)


class _SklearnRegressor(Model):
    """One-hot-flatten wrapper around an sklearn regressor (sklearn_models.py:12-35), CPU, third-party."""

    def __init__(self, estimator, alphabet: str, name: str):
        super().__init__(name)
        self.estimator, self.alphabet = estimator, alphabet

    def _features(self, sequences):
        idx = s_utils.encode_sequences(list(sequences), self.alphabet)
        out = np.zeros(idx.shape + (len(self.alphabet),), dtype=np.float32)
        np.put_along_axis(out, idx[..., None].astype(np.int64), 1.0, axis=2)
        return out.reshape(len(idx), -1)

    def train(self, sequences, labels):
        self.estimator.fit(self._features(sequences), np.asarray(labels, dtype=np.float64))

    def _fitness_function(self, sequences):
        return self.estimator.predict(self._features(sequences))


class DynaPPOEnsemble(Model):
    """Ensemble of heterogeneous models; only members whose held-out r^2 passes the threshold vote
    (dyna_ppo.py:32-130).  Default members: the B200 MLP(200) and CNN(32,100) plus the reference's sklearn
    regressors (the Keras GlobalEpistasisModel member is out of this round's scope)."""

    def __init__(self, seq_len: int, alphabet: str, r_squared_threshold: float = 0.5,
                 models: Optional[List[Model]] = None):
        super().__init__(name="DynaPPOEnsemble")
        if models is None:
            import sklearn.ensemble, sklearn.gaussian_process, sklearn.linear_model, sklearn.neighbors, sklearn.tree

            from flexs_b200.baselines.models import CNN, MLP

            models = [
                MLP(seq_len, 200, alphabet), CNN(seq_len, 32, 100, alphabet),
                _SklearnRegressor(sklearn.linear_model.LinearRegression(), alphabet, "linear_regression"),
                _SklearnRegressor(sklearn.ensemble.RandomForestRegressor(), alphabet, "random_forest"),
                _SklearnRegressor(sklearn.neighbors.KNeighborsRegressor(), alphabet, "nearest_neighbors"),
                _SklearnRegressor(sklearn.linear_model.Lasso(), alphabet, "lasso"),
                _SklearnRegressor(sklearn.linear_model.BayesianRidge(), alphabet, "bayesian_ridge"),
                _SklearnRegressor(sklearn.gaussian_process.GaussianProcessRegressor(), alphabet, "gaussian_process"),
                _SklearnRegressor(sklearn.ensemble.GradientBoostingRegressor(), alphabet, "gradient_boosting"),
                _SklearnRegressor(sklearn.tree.ExtraTreeRegressor(), alphabet, "extra_trees"),
            ]
        self.models = models
        self.r_squared_vals = np.ones(len(models))
        self.r_squared_threshold = r_squared_threshold

    def train(self, sequences, labels):
        if len(sequences) < 10:
            return
        seqs, labs = np.array(sequences), np.array(labels)
        perm = np.random.permutation(len(seqs))
        n_test = int(np.ceil(0.25 * len(seqs)))
        test, train = perm[:n_test], perm[n_test:]
        for member in self.models:
            member.train(seqs[train], labs[train])
        self.r_squared_vals = []
        for member in self.models:
            preds = np.asarray(member.get_fitness(seqs[test]), dtype=np.float64)
            if (preds[0] == preds).all() or (labs[test][0] == labs[test]).all():
                self.r_squared_vals.append(0)
            else:
                again = np.asarray(member.get_fitness(seqs[test]), dtype=np.float64)  # the reference queries twice
                self.r_squared_vals.append(np.corrcoef(labs[test].astype(np.float64), again)[0, 1] ** 2)

    def _fitness_function(self, sequences):
        passing = [m for m, r2 in zip(self.models, self.r_squared_vals) if r2 >= self.r_squared_threshold]
        if not passing:
            return self.models[int(np.argmax(self.r_squared_vals))].get_fitness(sequences)
        return np.mean([m.get_fitness(sequences) for m in passing], axis=0)


class _ConstructiveEnv:
    """Batched constructive environment (environments/dyna_ppo.py:13-163)."""

    def __init__(self, alphabet: str, seq_length: int, model: Model, landscape: Landscape, batch_size: int):
        self.alphabet, self.seq_length, self.batch_size = alphabet, seq_length, batch_size
        self.model, self.landscape = model, landscape
        self.fitness_model_is_gt = False
        self.all_seqs: Dict[str, float] = {}
        self.lam = 0.1
        self.reset()

    def reset(self):
        self.partial_seq_len = 0
        self.states = np.zeros((self.batch_size, self.seq_length, len(self.alphabet) + 1), dtype=np.float32)
        self.states[:, :, -1] = 1
        return self.states.copy()

    def sequence_density(self, seq: str) -> float:
        dens = 0.0
        for other, fit in self.all_seqs.items():
            dist = bounded_edit_distance(other, seq, 2)
            if dist != 0 and dist <= 2:
                dens += fit / dist
        return dens

    def step(self, actions: np.ndarray):
        """Returns (observation, reward, done, complete_sequences or None)."""
        actions = np.asarray(actions).reshape(-1)
        self.states[:, self.partial_seq_len, -1] = 0
        self.states[np.arange(self.batch_size), self.partial_seq_len, actions] = 1
        self.partial_seq_len += 1
        if self.partial_seq_len < self.seq_length - 1:
            return self.states.copy(), np.zeros(self.batch_size, dtype=np.float32), False, None
        idx = np.argmax(self.states[:, :, :-1], axis=2).astype(np.uint8)  # unfilled last row -> alphabet[0]
        complete = list(s_utils.decode_indices(idx, self.alphabet))
        source = self.landscape if self.fitness_model_is_gt else self.model
        fitnesses = np.asarray(source.get_fitness(complete), dtype=np.float64)     # HOT CALL
        self.all_seqs.update(zip(complete, fitnesses))
        rewards = np.array([f - self.lam * self.sequence_density(s) for s, f in zip(complete, fitnesses)], dtype=np.float32)
        return self.states.copy(), rewards, True, complete


class _PPOAgent:
    def __init__(self, obs_dim: int, n_actions: int, lr: float = 1e-5, epochs: int = 10, clip: float = 0.2,
                 gamma: float = 0.99, lam: float = 0.95):
        self.dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        self.actor = nn.Sequential(nn.Linear(obs_dim, 128), nn.ReLU(), nn.Linear(128, n_actions)).to(self.dev)
        self.critic = nn.Sequential(nn.Linear(obs_dim, 128), nn.ReLU(), nn.Linear(128, 1)).to(self.dev)
        self.opt = torch.optim.Adam(list(self.actor.parameters()) + list(self.critic.parameters()), lr=lr)
        self.epochs, self.clip, self.gamma, self.lam = epochs, clip, gamma, lam

    def act(self, obs: np.ndarray):
        with torch.no_grad():
            x = torch.as_tensor(obs.reshape(len(obs), -1), device=self.dev)
            dist = torch.distributions.Categorical(logits=self.actor(x))
            a = dist.sample()
            return a.cpu().numpy(), dist.log_prob(a).cpu().numpy(), self.critic(x)[:, 0].cpu().numpy()

    def train(self, episodes: List[dict]):
        if not episodes:
            return
        obs, act, logp, adv, ret = [], [], [], [], []
        for ep in episodes:  # ep arrays are [T, E, ...]
            v = np.concatenate([ep["val"], np.zeros((1,) + ep["val"].shape[1:], dtype=np.float32)])
            gae = np.zeros_like(ep["rew"][0])
            advs = np.zeros_like(ep["rew"])
            for t in reversed(range(len(ep["rew"]))):
                delta = ep["rew"][t] + self.gamma * v[t + 1] - v[t]
                gae = delta + self.gamma * self.lam * gae
                advs[t] = gae
            obs.append(ep["obs"].reshape(-1, ep["obs"].shape[-2] * ep["obs"].shape[-1]))
            act.append(ep["act"].reshape(-1)); logp.append(ep["logp"].reshape(-1))
            adv.append(advs.reshape(-1)); ret.append((advs + ep["val"]).reshape(-1))
        obs = torch.as_tensor(np.concatenate(obs), device=self.dev)
        act = torch.as_tensor(np.concatenate(act), device=self.dev, dtype=torch.long)
        logp = torch.as_tensor(np.concatenate(logp), device=self.dev)
        adv = torch.as_tensor(np.concatenate(adv), device=self.dev)
        ret = torch.as_tensor(np.concatenate(ret), device=self.dev)
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        for _ in range(self.epochs):
            dist = torch.distributions.Categorical(logits=self.actor(obs))
            ratio = torch.exp(dist.log_prob(act) - logp)
            pol = -torch.min(ratio * adv, torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * adv).mean()
            val = 0.5 * (self.critic(obs)[:, 0] - ret).pow(2).mean()
            self.opt.zero_grad()
            (pol + val).backward()
            self.opt.step()


class DynaPPO(Explorer):
    """Model-based PPO sequence designer (constructive variant)."""

    def __init__(
        self,
        landscape: Landscape,
        rounds: int,
        sequences_batch_size: int,
        model_queries_per_batch: int,
        starting_sequence: str,
        alphabet: str,
        log_file: Optional[str] = None,
        model: Optional[Model] = None,
        num_experiment_rounds: int = 10,
        num_model_rounds: int = 1,
        env_batch_size: int = 4,
    ):
        """
        Args:
            num_experiment_rounds: experiment-based rounds (only used in the explorer's name, as in the reference).
            num_model_rounds: model-based policy-update rounds per proposal round.
            env_batch_size: episodes run in parallel — with a B200 surrogate raise this to thousands: every
                episode end is one fused-kernel launch whatever the batch.
        """
        if model is None:
            model = DynaPPOEnsemble(len(starting_sequence), alphabet)
            model.train(s_utils.generate_random_sequences(len(starting_sequence), 10, alphabet), [0] * 10)
        super().__init__(model, f"DynaPPO_Agent_{num_experiment_rounds}_{num_model_rounds}", rounds,
                         sequences_batch_size, model_queries_per_batch, starting_sequence, log_file)
        self.alphabet = alphabet
        self.num_experiment_rounds = num_experiment_rounds
        self.num_model_rounds = num_model_rounds
        self.env_batch_size = env_batch_size
        self.env = _ConstructiveEnv(alphabet, len(starting_sequence), model, landscape, env_batch_size)
        self.agent = _PPOAgent(len(starting_sequence) * (len(alphabet) + 1), len(alphabet))

    def _run_episode(self, new_seqs: Optional[dict]) -> dict:
        obs = self.env.reset()
        rec = {"obs": [], "act": [], "logp": [], "val": [], "rew": []}
        done = False
        while not done:
            a, logp, v = self.agent.act(obs)
            nxt, rew, done, complete = self.env.step(a)
            for key, val in zip(rec, (obs, a, logp, v, rew)):
                rec[key].append(val)
            obs = nxt
        if new_seqs is not None and complete is not None:
            for seq in complete:
                new_seqs[seq] = self.env.all_seqs[seq]
        return {k: np.asarray(v, dtype=np.float32 if k != "act" else np.int64) for k, v in rec.items()}

    def propose_sequences(self, measured_sequences_data: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """Return the ``sequences_batch_size`` best unmeasured sequences found in the model-based rounds."""
        # experiment-based round: rewards from the ground truth, budget = one proposal batch (:284-297)
        self.env.fitness_model_is_gt = True
        start = self.env.landscape.cost
        episodes = []
        while self.env.landscape.cost - start < self.sequences_batch_size:
            episodes.append(self._run_episode(None))
        self.agent.train(episodes)

        # model-based rounds (:299-307)
        found: Dict[str, float] = {}
        self.env.fitness_model_is_gt = False
        start = self.model.cost
        for _ in range(self.num_model_rounds):
            if self.model.cost - start >= self.model_queries_per_batch:
                break
            round_start = self.model.cost
            episodes = []
            while self.model.cost - round_start < int(self.model_queries_per_batch / self.num_model_rounds):
                episodes.append(self._run_episode(found))
            self.agent.train(episodes)

        measured = set(measured_sequences_data["sequence"])
        found = {s: f for s, f in found.items() if s not in measured}
        new_seqs = np.array(list(found.keys()))
        preds = np.array(list(found.values()))
        order = np.argsort(preds)[::-1][: self.sequences_batch_size]
        return new_seqs[order], preds[order]
