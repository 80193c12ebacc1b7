"""Run where TensorFlow/Keras IS installed (not possible in the authoring container): builds the reference's
CNN / MLP exactly as flexs/baselines/models/cnn.py:23-54 and mlp.py:21-31 do, and dumps
``get_weights()`` + ``predict()`` on fixed sequences to an .npz.  tests/test_gpu_parity.py picks the file up
from tests/golden/keras_*.npz when present, which turns "parity unpinned" into pinned for that shape.

    python tools/export_keras_golden.py --out tests/golden/keras_cnn_100x4.npz --seq-len 100 --alphabet TGCA
"""
import argparse

import numpy as np


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
    args = ap.parse_args()
    import tensorflow as tf  # noqa: F401  (must exist)

    L, A = args.seq_len, len(args.alphabet)
    if args.kind == "cnn":
        layers = tf.keras.layers
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
        layers = tf.keras.layers
        model = tf.keras.models.Sequential([
            layers.Flatten(input_shape=(L, A)),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(args.hidden_size, activation="relu"),
            layers.Dense(1),
        ])
    model.compile(loss="MSE", optimizer="adam", metrics=["mse"])
    rng = np.random.default_rng(0)
    # perturb the zero biases so the bias paths are exercised
    ws = [w + (rng.normal(0, 0.1, size=w.shape).astype(np.float32) if w.ndim == 1 else 0) for w in model.get_weights()]
    model.set_weights(ws)
    idx = rng.integers(0, A, size=(args.n, L), dtype=np.uint8)
    onehot = np.eye(A, dtype=np.float32)[idx]
    y = model.predict(onehot, batch_size=256).squeeze(axis=1)
    np.savez_compressed(args.out, idx=idx, y=np.nan_to_num(y), kind=args.kind, alphabet=args.alphabet,
                        kernel_size=args.kernel_size, **{f"w{i}": w for i, w in enumerate(model.get_weights())})
    print("wrote", args.out, "tensorflow", tf.__version__)


if __name__ == "__main__":
    main()
