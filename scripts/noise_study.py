"""How far is each class of tensor from fp64 -- ours vs the reference's own fp32 arithmetic -- over several seeds?
    SG2_PROMO_TAPS=<n> python scripts/noise_study.py [B] [seeds...]
Prints, per class, the worst and the median relative error of ours and of the reference fp32 draws.  Used to choose the promotion
interval of the fp32-class convolution (profiles/r2j_noise_study.txt): a setting is acceptable when ours sits inside the
reference's own spread."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fullwidth_common import evaluate, klass      # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    seeds = [int(s) for s in sys.argv[2:]] or [3, 4, 5]
    agg = {}
    for seed in seeds:
        rows, _ = evaluate(B, seed=seed, ref_draws=4)
        for k, e, draws, _ in rows:
            a = agg.setdefault(klass(k), dict(ours=[], ref=[]))
            a['ours'].append(e)
            a['ref'].extend(draws)
    med = lambda v: sorted(v)[len(v) // 2]
    print(f'promo_taps={os.environ.get("SG2_PROMO_TAPS", "default")} B={B} seeds={seeds}')
    for kl, a in sorted(agg.items()):
        print(f'   {kl[0]:12s} {kl[1]:6s} ours worst {max(a["ours"]):.2e} median {med(a["ours"]):.2e} | reference fp32 worst {max(a["ref"]):.2e} '
              f'median {med(a["ref"]):.2e}   (n = {len(a["ours"])} tensors)', flush=True)


if __name__ == '__main__':
    main()
