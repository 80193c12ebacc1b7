"""Dump real-Keras weights + predictions of the reference's surrogates as golden vectors (pins the float parity).

Run where TensorFlow/Keras IS installed (it is not in the authoring container).  ONE command per shape:

    python tools/export_keras_golden.py --out tests/golden/keras_cnn_100x4.npz  --seq-len 100 --alphabet TGCA
    python tools/export_keras_golden.py --out tests/golden/keras_cnn_237x20.npz --seq-len 237 --alphabet ILVAGMFYWEDQNHCRKSTP
    python tools/export_keras_golden.py --out tests/golden/keras_mlp_8x4.npz    --seq-len 8   --alphabet TGCA --kind mlp

The model is built exactly as flexs/baselines/models/cnn.py:23-54 / mlp.py:21-31 build it and evaluated as
keras_model.py:69-79 evaluates it (float32 one-hot, predict(batch_size=256), squeeze, nan_to_num).  Commit the .npz:
tests/test_gpu_parity.py::test_real_keras_golden_if_present (GPU) and
tests/test_oracle_golden.py::test_keras_golden_files_match_oracle (CPU) pick up every tests/golden/keras_*.npz through
tests/golden/keras_loader.py, and "parity unpinned" becomes pinned for those shapes.

``--backend torch`` writes the same file from torch-CPU layers with the Keras semantics (what this container can run);
it exists so that the file format and both consumers are exercised here — such a file does NOT pin anything and must not
be committed as keras_*.npz (the loader records the backend; the tests refuse non-tensorflow files in tests/golden/).
"""
import argparse

import numpy as np


def build_tf(args, L, A):
    import tensorflow as tf  # noqa: F401  (must exist)

    layers = tf.keras.layers
    if args.kind == "cnn":
        model = tf.keras.models.Sequential([
            layers.Conv1D(args.num_filters, args.kernel_size, padding="valid", activation="relu", strides=1, input_shape=(L, A)),
            layers.Conv1D(args.num_filters, args.kernel_size, padding="same", activation="relu", strides=1),
            layers.MaxPooling1D(1),
            layers.Conv1D(args.num_filters, A - 1, padding="same", activation="relu", strides=1),
            layers.GlobalMaxPooling1D(),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dropout(0.25),
            layers.Dense(1),
        ])
    else:
        model = tf.keras.models.Sequential([
            layers.Flatten(input_shape=(L, A)),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(1),
        ])
    model.compile(loss="MSE", optimizer="adam", metrics=["mse"])
    return model, "tensorflow " + tf.__version__


class TorchAsKeras:
    """get_weights / set_weights / predict of the same stacks on torch-CPU (Keras layouts at the boundary)."""

    def __init__(self, args, L, A, rng):
        F, H, K = args.num_filters, args.hidden_size, args.kernel_size

        def glorot(shape):
            rec = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
            lim = np.sqrt(6.0 / (rec * shape[-2] + rec * shape[-1]))
            return rng.uniform(-lim, lim, size=shape).astype(np.float32)

        if args.kind == "cnn":
            shapes = [(K, A, F), (F,), (K, F, F), (F,), (A - 1, F, F), (F,), (F, H), (H,), (H, H), (H,), (H, 1), (1,)]
        else:
            shapes = [(L * A, H), (H,), (H, H), (H,), (H, H), (H,), (H, 1), (1,)]
        self.kind, self.K, self.K3 = args.kind, K, A - 1
        self.w = [glorot(s) if len(s) > 1 else np.zeros(s, np.float32) for s in shapes]

    def get_weights(self):
        return [w.copy() for w in self.w]

    def set_weights(self, ws):
        self.w = [np.asarray(w, np.float32) for w in ws]

    def predict(self, onehot, batch_size=256):
        import torch
        import torch.nn.functional as Fn

        w = [torch.from_numpy(a) for a in self.w]
        outs = []
        with torch.no_grad():
            for i in range(0, len(onehot), batch_size):
                x = torch.from_numpy(onehot[i: i + batch_size])
                if self.kind == "cnn":
                    def same(h, k):
                        left = (k - 1) // 2
                        return Fn.pad(h, (left, (k - 1) - left))
                    h = x.permute(0, 2, 1)
                    h = Fn.relu(Fn.conv1d(h, w[0].permute(2, 1, 0).contiguous(), w[1]))
                    h = Fn.relu(Fn.conv1d(same(h, self.K), w[2].permute(2, 1, 0).contiguous(), w[3]))
                    h = Fn.relu(Fn.conv1d(same(h, self.K3), w[4].permute(2, 1, 0).contiguous(), w[5]))
                    h = h.amax(dim=2)
                    rest = w[6:]
                else:
                    h = x.reshape(len(x), -1)
                    rest = w
                for j in range(0, len(rest) - 2, 2):
                    h = Fn.relu(h @ rest[j] + rest[j + 1])
                outs.append((h @ rest[-2] + rest[-1]).numpy())
        return np.concatenate(outs)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--seq-len", type=int, default=100)
    ap.add_argument("--alphabet", default="TGCA")
    ap.add_argument("--num-filters", type=int, default=32)
    ap.add_argument("--hidden-size", type=int, default=100)
    ap.add_argument("--kernel-size", type=int, default=5)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--kind", default="cnn", choices=["cnn", "mlp"])
    ap.add_argument("--backend", default="tensorflow", choices=["tensorflow", "torch"])
    args = ap.parse_args()
    L, A = args.seq_len, len(args.alphabet)
    rng = np.random.default_rng(0)
    if args.backend == "tensorflow":
        model, backend = build_tf(args, L, A)
    else:
        model, backend = TorchAsKeras(args, L, A, rng), "torch"
    # perturb the zero biases so the bias paths are exercised
    ws = [w + (rng.normal(0, 0.1, size=w.shape).astype(np.float32) if w.ndim == 1 else 0) for w in model.get_weights()]
    model.set_weights(ws)
    idx = rng.integers(0, A, size=(args.n, L), dtype=np.uint8)
    onehot = np.eye(A, dtype=np.float32)[idx]
    y = np.asarray(model.predict(onehot, batch_size=256)).squeeze(axis=1)
    np.savez_compressed(args.out, idx=idx, y=np.nan_to_num(y).astype(np.float32), kind=args.kind, alphabet=args.alphabet,
                        kernel_size=args.kernel_size, backend=backend.split()[0],
                        **{f"w{i}": np.asarray(w, np.float32) for i, w in enumerate(model.get_weights())})
    print("wrote", args.out, "backend", backend)


if __name__ == "__main__":
    main()
