class Engine:  # name only
    pass
