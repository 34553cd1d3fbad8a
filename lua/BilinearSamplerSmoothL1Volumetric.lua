--[[
   nn.BilinearSamplerSmoothL1Volumetric -- the training step around a plain nn.BilinearSamplerVolumetric in ONE kernel
   (torch.CudaTensor only; single-channel volumes): for a network that emits a dense field B x D x H x W x 4 (or x 3),

       output = sampler:forward({inputImages, grids});  err = nn.SmoothL1Criterion():forward(output, target)
       sampler:backward({inputImages, grids}, criterion:backward(output, target))            -- train.lua:123-126, model.lua:34

   forwardBackward({inputImages, grids}, target)  ->  err, {gradInputImages, gradGrids}
   gradGrids has the bits of the module chain; self.lossPerSample, self.iouCounts (B x 2), self.output (keepOutput) as in
   nn.CageWarpSmoothL1Volumetric.  Waits for its stream (it returns a Lua number, like criterion:forward).
]]
local Step, parent = torch.class('nn.BilinearSamplerSmoothL1Volumetric', 'nn.Module')

local ONLY_GRID, ZERO_GRADVOL = 1, 4

function Step:__init(sizeAverage, onlyGrid, keepOutput)
   parent.__init(self)
   self.sizeAverage = (sizeAverage == nil) and true or sizeAverage
   self.onlyGrid = (onlyGrid == nil) and true or onlyGrid
   self.keepOutput = keepOutput or false
   self.gradInput = {}
   self.stats = torch.DoubleTensor()
end

function Step:forwardBackward(input, target)
   local images, grids = input[1], input[2]
   assert(images:nDimension() == 5 and images:isContiguous() and images:size(5) == 1)
   assert(grids:nDimension() == 5 and grids:size(1) == images:size(1) and (grids:size(5) == 4 or grids:size(5) == 3))
   assert(target:nDimension() == 5 and target:size(5) == 1)
   for d = 1, 4 do assert(target:size(d) == grids:size(d)) end
   local B, n = images:size(1), target:nElement()
   self.workspace = self.workspace or images.new()
   self.gradInput[1] = self.gradInput[1] or images.new()
   self.gradInput[2] = (self.gradInput[2] or images.new()):resizeAs(grids)
   if self.onlyGrid then self.gradInput[1]:resize(0) else self.gradInput[1]:resizeAs(images) end
   if torch.type(self.output) ~= torch.type(images) then self.output = images.new() end
   if self.keepOutput then self.output:resizeAs(target) else self.output:resize(0) end
   -- nn.Module:type() (model:cuda()) converts every tensor field: `stats` must stay a HOST DoubleTensor, `output` follows the input
   if torch.type(self.stats) ~= 'torch.DoubleTensor' then self.stats = torch.DoubleTensor() end
   self.stats:resize(B, 3)
   local flags = ZERO_GRADVOL + (self.onlyGrid and ONLY_GRID or 0)
   local gradScale = self.sizeAverage and (1 / n) or 1
   images.nn.BilinearSamplerSmoothL1Volumetric_forwardBackward(self, images, grids, target:contiguous(), self.output,
                                                               self.gradInput[1], self.gradInput[2], self.workspace, self.stats,
                                                               gradScale, flags)
   self.lossPerSample = self.stats:select(2, 1)
   self.iouCounts = self.stats:narrow(2, 2, 2)
   local err = self.lossPerSample:sum()
   if self.sizeAverage then err = err / n end
   self.loss = err
   return err, self.gradInput
end
