def quantize_tensor(x, num_bits=8):
    raise NotImplementedError('off the bottleneck path')


def dequantize_tensor(q_x):
    raise NotImplementedError('off the bottleneck path')
