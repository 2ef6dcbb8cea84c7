def PRNGKey(seed):
    return int(seed)


def split(key, num=2):
    return tuple(key * 7919 + i + 1 for i in range(num))
