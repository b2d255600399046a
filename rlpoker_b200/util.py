"""Host sampler with the reference's exact semantics (rlpoker/util.py:6-27): restrict to the available
actions, normalise, then ONE numpy ``RandomState.choice`` draw from the global legacy RNG.  (The device
lanes restate the same sampler on a Philox stream; see csrc/sampled.cu.)"""
import numpy as np


def sample_action(strategy, available_actions=None):
    actions = [a for a in strategy if available_actions is None or a in available_actions]
    probs = np.array([strategy[a] for a in actions])
    assert np.sum(probs) > 0.0, "Oops: {}, {}".format(probs, actions)
    probs = probs / np.sum(probs)
    return actions[np.random.choice(len(actions), p=probs)]
