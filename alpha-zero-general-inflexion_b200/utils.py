"""Small helpers with the reference's names (utils.py:1-40): ``AverageMeter`` and ``dotdict``."""


class AverageMeter:
    """Weighted running mean of a scalar; prints as the mean in ``%.2e`` (the loss read-outs of NNetWrapper.train).
    Attributes as in the reference: ``val`` (last value), ``sum``, ``count``, ``avg``."""

    __slots__ = ("val", "sum", "count")

    def __init__(self):
        self.val, self.sum, self.count = 0.0, 0.0, 0

    @property
    def avg(self) -> float:
        return self.sum / self.count if self.count else 0.0

    def update(self, val, n: int = 1) -> None:
        self.val = val
        self.sum += n * val
        self.count += n

    def __repr__(self) -> str:
        return format(self.avg, ".2e")


class dotdict(dict):
    """dict with attribute access (``args.numMCTSSims``); pickles as a list of (key, value) pairs like the
    reference's, so argument objects saved by either implementation load in the other (utils.py:23-40)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None          # lets hasattr()/getattr(default) behave

    def __getstate__(self):
        return list(self.items())

    def __setstate__(self, pairs):
        self.update(pairs)
