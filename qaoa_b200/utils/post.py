"""Classical post-processing of sampled bit-strings by greedy bit flips (reference qaoa/utils/post.py:5-77).
`instance` needs `problem`, `flipper` (BitFlip) and `stat` (reset / add_sample / get_CVaR / get_Variance); strings are in
the sampler's printing order (qubit 0 last).  Host-only."""
import statistics as _stat

import numpy as np


def post_processing(instance, samples, K=5):
    """Boosts every sampled string, refills instance.stat with the boosted strings' costs and returns their histogram."""
    instance.stat.reset()
    if isinstance(samples, str):
        samples = [samples]
    boosted_hist = {}
    for string in samples:
        better = instance.flipper.boost_samples(problem=instance.problem, string=string, K=K)
        count = samples[string] if isinstance(samples, dict) else 1
        instance.stat.add_sample(instance.problem.cost(better[::-1]), count, better[::-1])
        boosted_hist[better] = boosted_hist.get(better, 0) + count
    return boosted_hist


def post_process_all_depths(instance, K=5):
    """100 independent post-processings of every depth's histogram: (means, variances) of the resulting -CVaR per depth."""
    means, variances = [], []
    for depth, hist in instance.samplecount_hists.items():
        if not isinstance(hist, dict):
            raise TypeError
        values = []
        for _ in range(100):
            post_processing(instance=instance, samples=hist, K=K)
            values.append(-instance.stat.get_CVaR())
        means.append(_stat.mean(values))
        variances.append(_stat.variance(values))
    return (np.array(means), np.array(variances))
