#!/usr/bin/env python
"""Build per-class barcode prototype windows for the synthetic reads (SURVEY.md 8d,
"Barcode-branch coverage"): adapter signals that the preset's own demultiplexer network
(barcoding.py:51-70, demux-tetra-r4) assigns to BC1..BC4 with a high score.

Synthetic adapters sampled from the HMM emissions are noise to the classifier: every one
of them comes back as the decoy class, so the accept branch (barcoding.py:108-118), the
calibration bins and the four barcode slots of the count tensor never see data.  Here a
float32 torch restatement of the network (Keras LSTMCell equations) is
differentiated with respect to its input: gradient ascent on the class probability, under
the random distortions a planted adapter undergoes before it reaches the classifier
(random adapter length, boundary jitter of the Viterbi segmentation, pooled sample noise,
median/MAD normalisation, -1000 left padding: barcoding.py:77-101).

Output: poreplex_b200/presets/synth_barcode_prototypes.npz with
  short [4][180]  right-aligned prototypes for 172..178-sample adapters (bench-short preset;
                  shorter windows are mostly -1000 padding, which the network -- trained on
                  adapters of 260 samples and more -- calls decoy whatever they hold)
  stock [4][300]  prototypes for adapters of 270 samples and more (stock preset)
in units of synth.ADAPTER_LEVEL[1] pA about ADAPTER_LEVEL[0], values in [-AMP_LO, AMP_HI].  Realism is not the aim;
parity needs the CUDA path to match the oracle on inputs that reach every branch.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from poreplex_b200 import params  # noqa: E402

AMP_LO, AMP_HI = 3.0, 2.2   # prototypes live in [-AMP_LO, AMP_HI] normalised units: the pA range
                            # that stays adapter-like for the segmentation HMM (synth.ADAPTER_LEVEL)
PAD = -1000.0
SHORT_LENGTHS = (172, 178)   # adapter length range of 4000-sample synthetic reads (synth.py)


def _lstm(layer):
    """Keras LSTM weights in a torch.nn.LSTM (same gate order i|f|c|o; the kernels are stored
    transposed, the single Keras bias goes to bias_ih)."""
    H = layer.recurrent.shape[0]
    m = torch.nn.LSTM(layer.kernel.shape[0], H, batch_first=True)
    with torch.no_grad():
        m.weight_ih_l0.copy_(torch.tensor(layer.kernel.T.copy()))
        m.weight_hh_l0.copy_(torch.tensor(layer.recurrent.T.copy()))
        m.bias_ih_l0.copy_(torch.tensor(layer.bias))
        m.bias_hh_l0.zero_()
    for p_ in m.parameters():
        p_.requires_grad_(False)
    return m


class Classifier:
    def __init__(self, dm):
        self.fw, self.bw, self.l2 = _lstm(dm.fwd), _lstm(dm.bwd), _lstm(dm.l2)
        self.Wd = torch.tensor(np.asarray(dm.dense_kernel, np.float32))
        self.bd = torch.tensor(np.asarray(dm.dense_bias, np.float32))

    def logits(self, win):                      # win [B,300]
        x = win[:, :, None]
        hf, _ = self.fw(x)
        hb, _ = self.bw(torch.flip(x, dims=[1]))
        h1 = torch.cat([hf, torch.flip(hb, dims=[1])], dim=2)
        h2, _ = self.l2(h1)
        return h2[:, -1] @ self.Wd + self.bd


def normalise_and_pad(sig, trim=300):
    """barcoding.py:77-101 for one 1-D tensor (differentiable)."""
    if sig.shape[0] > trim:
        sig = sig[-trim:]
    med = sig.median()
    mad = (sig - med).abs().median()
    w = (sig - med) / torch.clamp(mad * 1.4826, min=0.01)
    if w.shape[0] < trim:
        w = torch.cat([w.new_full((trim - w.shape[0],), PAD), w])
    return w


def distorted_batch(proto, labels_of, lengths, g, per_proto):
    """proto [R][P] (R candidate prototypes, labels_of[r] = class 1..4) -> windows
    [R * per_proto][300] under the distortions a planted adapter meets, and their labels."""
    wins, labels = [], []
    P = proto.shape[1]
    for k in range(proto.shape[0]):
        for _ in range(per_proto):
            M = int(torch.randint(lengths[0], lengths[1] + 1, (1,), generator=g))
            body = proto[k, P - min(M, P):]
            if M > P:                            # stock: anything adapter-like in front
                body = torch.cat([0.9 * torch.randn(M - P, generator=g), body])
            # Measured on synthetic reads through the oracle: the Viterbi segmentation puts the
            # adapter boundaries on the planted sample for > 90 % of the reads and within one or
            # two samples for the rest, so boundary jitter is rare and small.
            r = float(torch.rand(1, generator=g))
            if r < 0.08:
                body = body[:-1]
            elif r < 0.16:                       # one poly(A)-level sample swallowed
                body = torch.cat([body, 4.5 + 0.4 * torch.randn(1, generator=g)])
            r = float(torch.rand(1, generator=g))
            if r < 0.08:                         # a leader-high sample in front
                body = torch.cat([5.5 + 0.6 * torch.randn(1, generator=g), body])
            elif r < 0.16:
                body = body[int(torch.randint(1, 3, (1,), generator=g)):]
            sd = 0.05 + 0.10 * float(torch.rand(1, generator=g))
            body = body + sd * torch.randn(body.shape[0], generator=g)
            wins.append(normalise_and_pad(body))
            labels.append(int(labels_of[k]))
    return torch.stack(wins), torch.tensor(labels)


def bounded(theta):
    """Prototype values in [-AMP_LO, AMP_HI] normalised units."""
    t = torch.tanh(theta)
    return torch.where(t < 0, AMP_LO * t, AMP_HI * t)


def optimise(clf, P, lengths, iters, per_proto, seed, restarts=1, classes=(1, 2, 3, 4), keep=None):
    """Gradient ascent on the class probability; `restarts` independent candidates per class,
    the one with the best median score on fresh distorted copies is kept.  `keep` [4][P]: an
    earlier result; a class keeps its old prototype unless the new one validates better."""
    g = torch.Generator().manual_seed(seed)
    C = len(classes)
    R = C * restarts
    labels_of = torch.tensor([classes[r % C] for r in range(R)])
    theta = (0.3 * torch.randn(R, P, generator=g)).requires_grad_(True)
    opt = torch.optim.Adam([theta], lr=0.08)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, iters, eta_min=0.005)

    def scores(protos, labs, per):
        win, lab = distorted_batch(protos, labs, lengths, g, per)
        lg = clf.logits(win)
        p = torch.softmax(lg, dim=1)
        return lg, lab, p[torch.arange(len(lab)), lab].reshape(len(labs), per)

    for it in range(iters):
        lg, lab, sc = scores(bounded(theta), labels_of, per_proto)
        loss = torch.nn.functional.cross_entropy(lg, lab)
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if it % 200 == 0 or it == iters - 1:
            med = sc.detach().median(dim=1).values.reshape(restarts, C)
            print('  iter %4d loss %.4f median score per class (best restart) %s' %
                  (it, float(loss.detach()), med.max(dim=0).values.numpy().round(4)), flush=True)
    with torch.no_grad():
        # stored as optimised, i.e. inside [-AMP_LO, AMP_HI]: the generator maps a unit to a fixed
        # number of pA, and the window normalisation (barcoding.py:77-81) removes the affine part
        _, _, sc = scores(bounded(theta), labels_of, 64)
        med = sc.median(dim=1).values.reshape(restarts, C)
        best = med.argmax(dim=0)
        out = np.zeros((4, P), np.float32) if keep is None else np.array(keep, np.float32)
        old = np.zeros(4)
        if keep is not None:
            _, _, sk = scores(torch.tensor(out), torch.arange(4) + 1, 64)
            old = sk.median(dim=1).values.numpy()
        for j, k in enumerate(classes):
            new = float(med[best[j], j])
            print('  class %d: validation median score %.4f (kept prototype: %.4f)' % (k, new, old[k - 1]))
            if keep is None or new > old[k - 1]:
                out[k - 1] = bounded(theta)[int(best[j]) * C + j].numpy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=3000)
    ap.add_argument('--per-class', type=int, default=12)
    ap.add_argument('--seed', type=int, default=20261017)
    ap.add_argument('--sets', default='short,stock')
    ap.add_argument('--restarts', type=int, default=4)
    ap.add_argument('--classes', default='1,2,3,4', help='short set: classes to (re)optimise')
    ap.add_argument('--out', default=os.path.join(params.PRESET_DIR, 'synth_barcode_prototypes.npz'))
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    preset = params.load_preset()
    dm = params.load_demux_model(preset['demultiplexing']['demux_model'])
    clf = Classifier(dm)
    classes = tuple(int(c) for c in a.classes.split(','))
    out = dict(np.load(a.out)) if os.path.exists(a.out) else {}
    if 'short' in a.sets:
        print('short prototypes (%d..%d pooled samples, bench-short preset)' % SHORT_LENGTHS)
        out['short'] = optimise(clf, SHORT_LENGTHS[1], SHORT_LENGTHS, a.iters, a.per_class, a.seed,
                                restarts=a.restarts, classes=classes,
                                keep=out.get('short') if len(classes) < 4 else None)
    if 'stock' in a.sets:
        print('stock prototypes (>= 270 pooled samples)')
        out['stock'] = optimise(clf, 300, (270, 330), min(a.iters, 600), a.per_class, a.seed + 1)
    np.savez(a.out, **out)
    print('wrote', a.out, {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
