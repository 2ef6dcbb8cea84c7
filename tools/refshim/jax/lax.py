def stop_gradient(x):
    return x


def reduce_precision(x, exponent_bits, mantissa_bits):
    return x
