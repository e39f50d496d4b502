"""Small helpers with the reference's names (utils.py:1-40): ``AverageMeter`` and ``dotdict``."""


class AverageMeter:
    """Running mean of a scalar (used for the loss read-outs of NNetWrapper.train)."""

    def __init__(self):
        self.val = self.avg = self.sum = 0.0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __repr__(self):
        return f"{self.avg:.2e}"


class dotdict(dict):
    """dict with attribute access; pickles as a list of pairs like the reference's (utils.py:23-40)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __getstate__(self):
        return list(self.items())

    def __setstate__(self, state):
        self.update(dict(state))
