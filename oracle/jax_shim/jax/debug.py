def callback(func, *args, **kwargs):
    return func(*args, **kwargs)
