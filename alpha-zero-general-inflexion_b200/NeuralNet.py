"""Abstract policy/value network interface (reference: NeuralNet.py:1-50), kept verbatim in shape so that
existing wrappers plug in.  ``predict_batch`` is the one addition: the batched device-side call the batched
MCTS uses instead of one ``predict`` per leaf."""


class NeuralNet:
    def __init__(self, game):
        pass

    def train(self, examples):
        """examples: list of (planes int64[4,7,7], pi list[343], v float)."""
        raise NotImplementedError

    def predict(self, board):
        """planes [4,7,7] -> (pi float32[343], v float32 ndarray of shape (1,))."""
        raise NotImplementedError

    def predict_batch(self, planes):
        """device tensor float32 [B,4,7,7] -> (pi float32 [B,343], v float32 [B]) on the same device."""
        raise NotImplementedError

    def save_checkpoint(self, folder, filename):
        raise NotImplementedError

    def load_checkpoint(self, folder, filename):
        raise NotImplementedError
