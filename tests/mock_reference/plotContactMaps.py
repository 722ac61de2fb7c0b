def plotContactMap(*args, **kwargs):
    return None
