from abc import ABCMeta, abstractmethod


class Trainer(metaclass=ABCMeta):
    """Same one-method contract as the reference's train/trainer.py:4-7."""

    @abstractmethod
    def train(self):
        ...
