"""ArrayInput (stoix/networks/inputs.py:7-12): identity input layer."""


class ArrayInput:
    def __call__(self, embedding):
        return embedding
