--[[
   nn.CageWarpSmoothL1Volumetric -- one training step of ALIGNet's warp head in ONE kernel (torch.CudaTensor only).

   forwardBackward({source, cage}, target) does, for single-channel volumes, what train.lua:123-126 does with
   nn.SmoothL1Criterion (model.lua:34) around the warp sub-graph of nn.CageWarpVolumetric:

       output = model:forward({source, cage});  err = criterion:forward(output, target)
       model:backward({source, cage}, criterion:backward(output, target))

   and returns  err, {gradSource, gradCage}.  Also left behind: self.lossPerSample (B), self.iouCounts (B x 2:
   intersection and union at threshold 0.5 per sample; util.lua:110-120 thresholds the same way but groups elements by
   flat index modulo B -- computeIOU on a kept output reproduces that number), self.output when keepOutput is set.
   Like criterion:forward, the call returns a Lua number and therefore waits for its stream.
]]
local Step, parent = torch.class('nn.CageWarpSmoothL1Volumetric', 'nn.Module')

local ONLY_GRID, ZERO_GRADVOL, FAST_ARITH = 1, 4, 0x40000

function Step:__init(sizeAverage, onlyGrid, keepOutput, fastArith)
   parent.__init(self)
   self.sizeAverage = (sizeAverage == nil) and true or sizeAverage     -- nn.SmoothL1Criterion's default
   self.onlyGrid = (onlyGrid == nil) and true or onlyGrid              -- the source is data (models/alignet_2d.lua:23-24)
   self.keepOutput = keepOutput or false
   self.fastArith = fastArith or false
   self.gradInput = {}
   self.stats = torch.DoubleTensor()
end

function Step:forwardBackward(input, target)
   local source, cage = input[1], input[2]
   assert(source:nDimension() == 5 and source:isContiguous() and source:size(5) == 1, 'source: contiguous B x D x H x W x 1 expected')
   assert(cage:nDimension() == 5 and cage:size(2) == 3 and cage:isContiguous(), 'cage: contiguous B x 3 x cD x cH x cW expected')
   assert(target:nDimension() == 5 and target:size(1) == source:size(1) and target:size(5) == 1)
   local B, n = source:size(1), target:nElement()
   self.workspace = self.workspace or source.new()
   self.gradInput[1] = self.gradInput[1] or source.new()
   self.gradInput[2] = (self.gradInput[2] or source.new()):resizeAs(cage)
   if self.onlyGrid then self.gradInput[1]:resize(0) else self.gradInput[1]:resizeAs(source) end
   if torch.type(self.output) ~= torch.type(source) then self.output = source.new() end
   if self.keepOutput then self.output:resizeAs(target) else self.output:resize(0) end
   -- nn.Module:type() (model:cuda()) converts every tensor field: `stats` must stay a HOST DoubleTensor, `output` follows the input
   if torch.type(self.stats) ~= 'torch.DoubleTensor' then self.stats = torch.DoubleTensor() end
   self.stats:resize(B, 3)
   local flags = ZERO_GRADVOL + (self.onlyGrid and ONLY_GRID or 0) + (self.fastArith and FAST_ARITH or 0)
   local gradScale = self.sizeAverage and (1 / n) or 1                 -- THNN: norm = sizeAverage ? 1. / nElement : 1.
   source.nn.CageWarpSmoothL1Volumetric_forwardBackward(self, source, cage, target, self.output, self.gradInput[1],
                                                        self.gradInput[2], self.workspace, self.stats, gradScale, flags)
   self.lossPerSample = self.stats:select(2, 1)
   self.iouCounts = self.stats:narrow(2, 2, 2)
   local err = self.lossPerSample:sum()
   if self.sizeAverage then err = err / n end
   self.loss = err
   return err, self.gradInput
end

-- mean over the batch of per-sample intersection / union
function Step:iou()
   return torch.cdiv(self.iouCounts:select(2, 1), self.iouCounts:select(2, 2)):mean()
end
