from functools import wraps


def coroutine(func):
    """Primes a generator so that the first `.send()` delivers a value (reference: src/coroutines/__init__.py)."""

    @wraps(func)
    def primed(*args, **kwargs):
        gen = func(*args, **kwargs)
        next(gen)
        return gen

    return primed
