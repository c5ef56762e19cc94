class Accelerator:  # name only
    pass
