def __getattr__(name):
    raise RuntimeError(f'matplotlib.pyplot.{name} is not available in the golden-generation harness')
