import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loans_b200.functions import FrameIngest
dev=torch.device('cuda')
raw=torch.randint(0,256,(64,384,512,3),dtype=torch.uint8,device=dev)
ing=FrameIngest(64,(384,512),(224,224))
for _ in range(3): out=ing(raw)
torch.cuda.synchronize()
