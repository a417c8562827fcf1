"""Stream sharding across the GPUs of one box: streams are independent (no exchange step, hence no collective), so rank r of N
simply owns the streams s with s % N == r.  Used by bench.py; covered on CPU with gloo by tests/test_sharding.py."""


def streams_of_rank(rank, world, total_streams):
    return list(range(rank, total_streams, world))


def stream_seed(stream_index, base=12345):
    """seed of the synthetic G1 sequence of a stream (distinct content per stream)"""
    return base + 7919 * stream_index
