"""tell/facenet on the B200 path: InceptionResnetV1 face embedder and the MTCNN P/R/O networks."""
from .inception_resnet_v1 import InceptionResnetV1
from .mtcnn import ONet, PNet, RNet

__all__ = ['InceptionResnetV1', 'PNet', 'RNet', 'ONet']
