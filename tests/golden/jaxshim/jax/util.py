def unzip2(xys):
    xs, ys = [], []
    for x, y in xys:
        xs.append(x); ys.append(y)
    return tuple(xs), tuple(ys)
