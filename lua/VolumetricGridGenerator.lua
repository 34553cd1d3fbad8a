--[[
   nn.VolumetricGridGenerator(depth, height, width) -- drop-in for stn3dtorch/VolumetricGridGenerator.lua.

   updateOutput(T): T is b x 4 x 4 (or 4 x 4); output[b][k][i][j] = T[b] * (z_k, y_i, x_j, 1) with
   z_k = -1 + (k-1)/(depth-1)*2 etc. -- the product the reference forms with torch.bmm over a batch-replicated base grid
   (lua:50-60).  Here one kernel writes it, the axis values evaluated in double and rounded like the Lua constructor
   (lua:17-19); updateGradInput is one deterministic reduction instead of zero() + baddbmm (lua:79-82).

   Serialised fields kept: depth, height, width, baseGrid (depth x height x width x 4), batchGrid.  batchGrid is a
   1 x depth x height x width x 4 VIEW of baseGrid and is never replicated over the batch any more.
]]
local Generator, parent = torch.class('nn.VolumetricGridGenerator', 'nn.Module')

function Generator:__init(depth, height, width)
   parent.__init(self)
   assert(depth > 1)
   assert(height > 1)
   assert(width > 1)
   self.depth, self.height, self.width = depth, height, width
   -- the base grid is the generator's own output for the identity transform
   local eye = torch.eye(4):view(1, 4, 4)
   self.baseGrid = torch.Tensor(1, depth, height, width, 4)
   eye.nn.VolumetricGridGenerator_updateOutput(self, eye, self.baseGrid)
   self.batchGrid = self.baseGrid
   self.baseGrid = self.baseGrid:select(1, 1)
end

local function matrices(t)
   if t:nDimension() == 2 then t = t:view(1, t:size(1), t:size(2)) end
   assert(t:nDimension() == 3 and t:size(2) == 4 and t:size(3) == 4, 'please input affine transform matrices (bx4x4)')
   return t
end

function Generator:updateOutput(transformMatrix)
   local single = transformMatrix:nDimension() == 2
   local T = matrices(transformMatrix):contiguous()
   self.output:resize(T:size(1), self.depth, self.height, self.width, 4)
   T.nn.VolumetricGridGenerator_updateOutput(self, T, self.output)
   if single then
      self.output = self.output:select(1, 1)
   end
   return self.output
end

function Generator:updateGradInput(transformMatrix, gradGrid)
   local single = transformMatrix:nDimension() == 2
   local T = matrices(transformMatrix)
   local gg = gradGrid
   if single then gg = gg:view(1, self.depth, self.height, self.width, 4) end
   gg = gg:contiguous()                       -- the reference's :view() (lua:77) needs it as well
   self.gradInput:resize(T:size(1), 4, 4)
   self._workspace = self._workspace or gg.new()   -- scratch for the two-stage reduction (CudaTensor path)
   gg.nn.VolumetricGridGenerator_updateGradInput(self, gg, self.gradInput, self._workspace)
   if single then
      self.gradInput = self.gradInput:select(1, 1)
   end
   return self.gradInput
end

function Generator:clearState()
   nn.utils.clear(self, '_workspace')
   return parent.clearState(self)
end
