class HTML:
    def __init__(self, data=None, **kw):
        self.data = data


def display(*a, **kw):
    pass
