--[[
   nn.AffineSamplerVolumetric(depth, height, width) -- nn.VolumetricGridGenerator + nn.BilinearSamplerVolumetric as ONE call.

   Replaces the `projector` of stn3dtorch/test.lua:15-25 and the loader-side augmentation pattern of dataset_2d.lua:285-311
   (lifted to 3-D):

       grid   = nn.VolumetricGridGenerator(depth, height, width):forward(T)        -- T: B x 4 x 4
       output = nn.BilinearSamplerVolumetric():forward({inputImages, grid})

   updateOutput({inputImages, T})  ->  B x depth x height x width x C      (inputImages B x D x H x W x C contiguous)
   updateGradInput({inputImages, T}, gradOutput)  ->  {gradInputImages, gradT}

   The B x depth x height x width x 4 grid (16 bytes per output voxel each way) never exists: the coordinates
   T[b] . (z, y, x, 1) are formed in registers with the grid generator's own operation order, so `output` has the bits
   of the two-module chain.  torch.FloatTensor (staged through the GPU) and torch.CudaTensor.
     self.onlyGrid   skip gradInputImages (returned empty)
]]
local Affine, parent = torch.class('nn.AffineSamplerVolumetric', 'nn.Module')

local ONLY_GRID, ZERO_GRADVOL = 1, 4

function Affine:__init(depth, height, width, onlyGrid)
   parent.__init(self)
   assert(depth > 1 and height > 1 and width > 1)          -- VolumetricGridGenerator.lua:6-8
   self.depth, self.height, self.width = depth, height, width
   self.onlyGrid = onlyGrid or false
   self.gradInput = {}
end

function Affine:check(input)
   local images, T = input[1], input[2]
   assert(images:nDimension() == 5 and images:isContiguous(), 'Input images have to be contiguous B x D x H x W x C')
   assert(T:nDimension() == 3 and T:size(1) == images:size(1) and T:size(2) == 4 and T:size(3) == 4,
          'please input affine transform matrices (bx4x4)')
end

function Affine:updateOutput(input)
   self:check(input)
   local images, T = input[1], input[2]:contiguous()
   if torch.type(self.output) ~= torch.type(images) then self.output = images.new() end
   self.output:resize(images:size(1), self.depth, self.height, self.width, images:size(5))
   images.nn.AffineSamplerVolumetric_updateOutput(self, images, T, self.output)
   return self.output
end

function Affine:updateGradInput(input, gradOutput)
   self:check(input)
   local images, T = input[1], input[2]:contiguous()
   local flags = ZERO_GRADVOL + (self.onlyGrid and ONLY_GRID or 0)
   self.gradInput[1] = self.gradInput[1] or images.new()
   if self.onlyGrid then self.gradInput[1]:resize(0) else self.gradInput[1]:resizeAs(images) end
   self.gradInput[2] = (self.gradInput[2] or images.new()):resizeAs(T)
   if torch.type(images) == 'torch.CudaTensor' then
      self.workspace = self.workspace or images.new()       -- resized by the native function
      images.nn.AffineSamplerVolumetric_updateGradInput(self, images, T, self.gradInput[1], self.gradInput[2], gradOutput,
                                                        self.workspace, flags)
   else
      images.nn.AffineSamplerVolumetric_updateGradInput(self, images, T, self.gradInput[1], self.gradInput[2], gradOutput, flags)
   end
   return self.gradInput
end
