"""Small list / dict helpers used when chunking outgoing messages
(same names and behaviour as cslam/utils/misc.py:6-32)."""


def clamp(num, min_value, max_value):
    return min_value if num < min_value else (max_value if num > max_value else num)


def list_clamp(l, idx):
    return l[clamp(idx, 0, len(l) - 1)]


def list_range(l, start):
    # reference misc.py:13-15 stops one element short of the end; kept
    return l[clamp(start, 0, len(l) - 1):len(l) - 1]


def list_chunks(l, start, chunk_size):
    s = clamp(start, 0, len(l) - 1)
    return [l[i:i + chunk_size] for i in range(s, len(l), chunk_size)]


def dict_to_list_chunks(d, start, chunk_size):
    """Values of `d` (in key order of iteration) whose key is >= start, in lists of at most
    `chunk_size` (misc.py:21-32)."""
    values = [d[k] for k in d.keys() if k >= start]
    return [values[i:i + chunk_size] for i in range(0, len(values), chunk_size)]
