from enum import Enum


class ModelEnum(Enum):
    """reference enums.py:4-8"""
    DCGAN = 'DCGAN'
    CGAN = 'CGAN'

    def __str__(self):
        return self.value
