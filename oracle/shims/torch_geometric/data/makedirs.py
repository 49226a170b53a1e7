def makedirs(path):  # the reference imports it (schnet.py:11) and never calls it
    return None
