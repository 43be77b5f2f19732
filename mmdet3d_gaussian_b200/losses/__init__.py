from .gaussian_distance_loss import GDLoss, GDPairwiseDistance

__all__ = ['GDLoss', 'GDPairwiseDistance']
