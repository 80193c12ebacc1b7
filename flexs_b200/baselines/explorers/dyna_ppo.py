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


class _DeviceRollouts:
    """The constructive environment AND the PPO agent with everything resident on the GPU:
    thousands of parallel episodes instead of the reference's ``env_batch_size=4``.

    * A state of the reference's environment is the one-hot matrix ``(L, A+1)`` with a mask channel
      (environments/dyna_ppo.py:62-75); here it is implicit in the residues chosen so far, and the first layer of the
      actor / value networks — ``Linear(L*(A+1), 128)`` applied to that one-hot — is maintained INCREMENTALLY: choosing
      residue ``a`` at position ``t`` changes one mask bit and one residue bit, i.e. ``h += W1[t, a] - W1[t, mask]``, one
      row gather per step instead of an ``L*(A+1)``-wide product (same numbers up to summation order).
    * The episode ends after ``L-1`` steps; the last position is never sampled and decodes to ``alphabet[0]``
      (reference quirk, :137, :144-147).  The complete sequences go to the surrogate as ``uint8[E, L]`` residue indices
      (``get_fitness_device``: no strings, no one-hot) or — in the experiment-based round — to the landscape.
    * reward = fitness - 0.1 * density with the density over everything seen so far, this batch included (:148-160),
      computed by ``flexs_edit_density_dev`` (K8).  ``all_seqs`` is a pair of device arrays kept duplicate-free with
      ``flexs_dedup_representatives_dev``.
    * PPO update: for training the first-layer pre-activations of ALL steps of an episode are an exclusive prefix sum of
      gathered rows (``cumsum``), so the update touches ``E*L*128`` numbers, not ``E*L*L*(A+1)*128``.  Hyper-parameters as
      the host agent (Adam 1e-5, 10 epochs, clip 0.2).
    """

    def __init__(self, alphabet: str, seq_length: int, model: Model, landscape: Landscape, batch_size: int, device,
                 lr: float = 1e-5, epochs: int = 10, clip: float = 0.2, gamma: float = 0.99, lam: float = 0.95,
                 seed: Optional[int] = None):
        self.alphabet, self.L, self.A, self.E = alphabet, seq_length, len(alphabet), batch_size
        self.model, self.landscape, self.dev = model, landscape, device
        self.fitness_model_is_gt = False
        self.lam_density = 0.1
        self.epochs, self.clip, self.gamma, self.lam = epochs, clip, gamma, lam
        g = torch.Generator(device=device)
        g.manual_seed(0 if seed is None else int(seed))
        self.gen = g
        obs_dim, hid = seq_length * (self.A + 1), 128

        def uniform(shape, fan_in):
            bound = 1.0 / np.sqrt(fan_in)
            return nn.Parameter((torch.rand(shape, device=device, generator=g) * 2 - 1) * bound)

        # first layers as [L, A+1, 128] tables (row (t, c) of the Linear weight's transpose), the rest as usual
        self.params = {
            "aW1": uniform((seq_length, self.A + 1, hid), obs_dim), "ab1": uniform((hid,), obs_dim),
            "aW2": uniform((hid, self.A), hid), "ab2": uniform((self.A,), hid),
            "cW1": uniform((seq_length, self.A + 1, hid), obs_dim), "cb1": uniform((hid,), obs_dim),
            "cW2": uniform((hid, 1), hid), "cb2": uniform((1,), hid),
        }
        self.opt = torch.optim.Adam(list(self.params.values()), lr=lr)
        self.seen_rows = torch.empty((0, seq_length), dtype=torch.uint8, device=device)
        self.seen_fit = torch.empty((0,), dtype=torch.float64, device=device)

    # -- environment ------------------------------------------------------------------------------
    def _score(self, rows):
        if self.fitness_model_is_gt:
            if hasattr(self.landscape, "get_fitness_device"):
                return torch.as_tensor(self.landscape.get_fitness_device(rows), device=self.dev).to(torch.float64)
            seqs = list(s_utils.decode_indices(rows.cpu().numpy(), self.alphabet))
            return torch.as_tensor(np.asarray(self.landscape.get_fitness(seqs), dtype=np.float64), device=self.dev)
        return self.model.get_fitness_device(rows).to(torch.float64)

    def _remember(self, rows, fit):
        """all_seqs.update(zip(complete, fitnesses)) (:152) on device arrays: new sequences are appended once."""
        from flexs_b200 import _native

        allrows = torch.cat([self.seen_rows, rows])
        rep = torch.empty(len(allrows), dtype=torch.int64, device=self.dev)
        work = torch.empty(_native.dedup_workspace_bytes(len(allrows)), dtype=torch.uint8, device=self.dev)
        with torch.cuda.device(self.dev):
            _native.dedup_representatives_dev(allrows.data_ptr(), len(allrows), self.L, rep.data_ptr(), work.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
        n0 = len(self.seen_rows)
        own = torch.arange(n0, len(allrows), device=self.dev)
        first = rep[n0:] == own
        self.seen_rows = torch.cat([self.seen_rows, rows[first]]).contiguous()
        self.seen_fit = torch.cat([self.seen_fit, fit[first]]).contiguous()

    def _density(self, rows):
        from flexs_b200 import _native

        out = torch.empty(len(rows), dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            _native.edit_density_dev(rows.data_ptr(), len(rows), self.seen_rows.data_ptr(), self.seen_fit.data_ptr(),
                                     len(self.seen_rows), self.L, 2, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out

    # -- acting -----------------------------------------------------------------------------------
    def run_episode(self):
        """One batch of ``E`` constructive episodes.  Returns the record PPO trains on and the complete sequences."""
        P, E, L, A = self.params, self.E, self.L, self.A
        T = L - 1
        actions = torch.zeros((E, L), dtype=torch.uint8, device=self.dev)     # last column stays 0: alphabet[0]
        logps = torch.empty((T, E), dtype=torch.float32, device=self.dev)
        values = torch.empty((T, E), dtype=torch.float32, device=self.dev)
        with torch.no_grad():
            ha = (P["ab1"] + P["aW1"][:, A, :].sum(dim=0)).expand(E, -1).clone()   # all positions masked
            hc = (P["cb1"] + P["cW1"][:, A, :].sum(dim=0)).expand(E, -1).clone()
            for t in range(T):
                logits = torch.relu(ha) @ P["aW2"] + P["ab2"]
                values[t] = (torch.relu(hc) @ P["cW2"] + P["cb2"])[:, 0]
                probs = torch.softmax(logits, dim=1)
                a = torch.multinomial(probs, 1, generator=self.gen)[:, 0]
                logps[t] = torch.log_softmax(logits, dim=1).gather(1, a[:, None])[:, 0]
                actions[:, t] = a.to(torch.uint8)
                ha += P["aW1"][t, a] - P["aW1"][t, A]
                hc += P["cW1"][t, a] - P["cW1"][t, A]
            rows = actions.contiguous()
            fit = self._score(rows)                                         # HOT CALL: one fused forward for E sequences
            self._remember(rows, fit)
            reward = (fit - self.lam_density * self._density(rows)).to(torch.float32)
        return {"act": actions[:, :T], "logp": logps, "val": values, "rew": reward}, rows, fit

    # -- PPO update -------------------------------------------------------------------------------
    def _hidden(self, W1, b1, act):
        """First-layer pre-activations of every step of every episode: [E, T, 128] via an exclusive prefix sum."""
        A, T = self.A, act.shape[1]
        pos = torch.arange(T, device=self.dev)
        delta = W1[pos[None, :], act.long()] - W1[pos, A][None, :, :]       # [E, T, 128]: what choosing a_t changes
        base = b1 + W1[:, A, :].sum(dim=0)
        return base + torch.cumsum(delta, dim=1) - delta                    # state BEFORE step t

    def train(self, episodes: List[dict]):
        if not episodes:
            return
        act = torch.cat([ep["act"] for ep in episodes])                     # [N, T]
        logp_old = torch.cat([ep["logp"] for ep in episodes], dim=1).T      # [N, T]
        val_old = torch.cat([ep["val"] for ep in episodes], dim=1).T
        rew = torch.cat([ep["rew"] for ep in episodes])                     # [N]: reward arrives with the last step
        N, T = act.shape
        adv = torch.zeros((N, T), dtype=torch.float32, device=self.dev)
        gae = torch.zeros(N, dtype=torch.float32, device=self.dev)
        for t in reversed(range(T)):
            nxt = val_old[:, t + 1] if t + 1 < T else torch.zeros(N, device=self.dev)
            delta = (rew if t == T - 1 else 0.0) + self.gamma * nxt - val_old[:, t]
            gae = delta + self.gamma * self.lam * gae
            adv[:, t] = gae
        ret = adv + val_old
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
        P = self.params
        chunk = max(1, (1 << 22) // max(1, T))        # bound the [n, T, 128] activations of one backward pass
        for _ in range(self.epochs):
            self.opt.zero_grad()
            for lo in range(0, N, chunk):
                sl = slice(lo, min(N, lo + chunk))
                logits = torch.relu(self._hidden(P["aW1"], P["ab1"], act[sl])) @ P["aW2"] + P["ab2"]
                logp = torch.log_softmax(logits, dim=2).gather(2, act[sl].long()[..., None])[..., 0]
                ratio = torch.exp(logp - logp_old[sl])
                pol = -torch.min(ratio * adv[sl], torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * adv[sl]).sum()
                v = (torch.relu(self._hidden(P["cW1"], P["cb1"], act[sl])) @ P["cW2"] + P["cb2"])[..., 0]
                val = 0.5 * (v - ret[sl]).pow(2).sum()
                ((pol + val) / (N * T)).backward()
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
        device_env: Optional[bool] = None,
        seed: Optional[int] = None,
    ):
        """
        Args:
            num_experiment_rounds: experiment-based rounds (only used in the explorer's name, as in the reference).
            num_model_rounds: model-based policy-update rounds per proposal round.
            env_batch_size: episodes run in parallel — with a B200 surrogate raise this to thousands: every
                episode end is one fused-kernel launch whatever the batch.
            device_env: environment, density penalty and PPO agent resident on the GPU (``None``: for
                ``env_batch_size >= 256`` with a B200 surrogate), see :class:`_DeviceRollouts`.
            seed: seed of the device agent's initialisation and sampling (the reference is unseeded).
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
        self.landscape = landscape
        self.device_env = device_env
        self._dev = None
        if self._use_device():
            index = getattr(model, "device", None)
            if index is None and hasattr(model, "models"):
                index = getattr(model.models[0], "device", 0)
            self._dev = _DeviceRollouts(alphabet, len(starting_sequence), model, landscape, env_batch_size,
                                        torch.device("cuda", int(index or 0)), seed=seed)
            return
        self.env = _ConstructiveEnv(alphabet, len(starting_sequence), model, landscape, env_batch_size)
        self.agent = _PPOAgent(len(starting_sequence) * (len(alphabet) + 1), len(alphabet))

    def _use_device(self) -> bool:
        if self.device_env is not None:
            return bool(self.device_env)
        return self.env_batch_size >= 256 and hasattr(self.model, "get_fitness_device")

    def _propose_device(self, measured_sequences_data: pd.DataFrame) -> Tuple[np.ndarray, np.ndarray]:
        """The reference's budget loops (dyna_ppo.py:284-319) around :class:`_DeviceRollouts`."""
        from flexs_b200 import _native

        d = self._dev
        d.fitness_model_is_gt = True
        start = self.landscape.cost
        episodes = []
        while self.landscape.cost - start < self.sequences_batch_size:
            cost_before = self.landscape.cost
            rec, _, _ = d.run_episode()
            if self.landscape.cost == cost_before:       # a landscape scored on the device charges itself
                self.landscape.cost += d.E
            episodes.append(rec)
        d.train(episodes)

        d.fitness_model_is_gt = False
        found_rows, found_fit = [], []
        start = self.model.cost
        for _ in range(self.num_model_rounds):
            if self.model.cost - start >= self.model_queries_per_batch:
                break
            round_start = self.model.cost
            episodes = []
            while self.model.cost - round_start < int(self.model_queries_per_batch / self.num_model_rounds):
                rec, rows, fit = d.run_episode()
                episodes.append(rec)
                found_rows.append(rows); found_fit.append(fit)
            d.train(episodes)
        if not found_rows:
            return np.array([], dtype=str), np.array([], dtype=np.float32)
        dev = d.dev
        rows, fit = torch.cat(found_rows).contiguous(), torch.cat(found_fit).to(torch.float32).contiguous()
        # `found` is a dict keyed by sequence, minus everything already measured (:309-314): measured rows in front,
        # keep the first occurrence of every sequence that does not start there
        measured = torch.from_numpy(s_utils.encode_sequences(list(measured_sequences_data["sequence"]), self.alphabet)).to(dev)
        allrows = torch.cat([measured, rows])
        rep = torch.empty(len(allrows), dtype=torch.int64, device=dev)
        work = torch.empty(_native.dedup_workspace_bytes(len(allrows)), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            _native.dedup_representatives_dev(allrows.data_ptr(), len(allrows), d.L, rep.data_ptr(), work.data_ptr(), stream)
        n0 = len(measured)
        keep = torch.nonzero(rep[n0:] == torch.arange(n0, len(allrows), device=dev)).reshape(-1)
        if len(keep) == 0:
            return np.array([], dtype=str), np.array([], dtype=np.float32)
        preds = fit[keep].contiguous()
        k = min(self.sequences_batch_size, len(keep))    # np.argsort(preds)[::-1][:B]: B items (unlike the others)
        top_s = torch.empty(k, dtype=torch.float32, device=dev)
        top_i = torch.empty(k, dtype=torch.int64, device=dev)
        swork = torch.empty(_native.topk_select_workspace_bytes(), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _native.topk_select_dev(preds.data_ptr(), len(keep), k, 0, 0, 0, False, top_s.data_ptr(), top_i.data_ptr(), 0, 0,
                                    swork.data_ptr(), stream)
        winners = rows[keep[top_i]].cpu().numpy()
        return s_utils.decode_indices(winners, self.alphabet), top_s.cpu().numpy()

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
        if self._dev is not None:
            return self._propose_device(measured_sequences_data)
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
