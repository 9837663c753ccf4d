def _noop(*a, **k):
    return None


plot = show = pcolormesh = xlim = ylim = figure = savefig = _noop
