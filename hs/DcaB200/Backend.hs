-- | The two compiled operators of the run loop for `Backend = GPU`.
--
-- `runStep` (src/SimRunner.hs:60-97) takes accFn1 / accFn2 as plain Haskell functions on
-- Accelerate arrays; `runSim` builds them with `runN bkend` (src/SimRunner.hs:211-212).
-- For the GPU backend they are these two functions instead: same types, same meaning,
-- executed by libdca_b200 (dca_acc_get_action / dca_acc_grid_step_train).  They are pure
-- functions of their arguments, so `unsafePerformIO` is sound here exactly as it is inside
-- accelerate-llvm-native's own `runN`.
--
-- Marshalling uses only `Data.Array.Accelerate` (toList / fromList), already a dependency
-- of the reference (dca.cabal:41-64); accelerate-io would make it zero-copy but is not in
-- the reference's dependency set.
--
-- NOT COMPILED IN THIS REPOSITORY (no GHC in the build image); see INTEGRATION.md.
module DcaB200.Backend
  ( gpuGetAction
  , gpuGridStepTrain
  , gpuFeatureRep
  , GpuOrder(..)
  ) where

import Base
import Data.Array.Accelerate (Array, DIM1, DIM3, Scalar, Z(..), (:.)(..))
import qualified Data.Array.Accelerate as A
import Data.Int
import Data.Word
import DcaB200.FFI
import Foreign.C.String (peekCString)
import Foreign.C.Types
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, withArray)
import Foreign.Ptr (nullPtr)
import Foreign.Storable (peek, poke)
import System.IO.Unsafe (unsafePerformIO)

-- | fp32 reduction order of the dots: the Accelerate Interpreter's right fold, or the
-- fixed tree of the throughput kernel (DESIGN.md, "Arithmetic orders").
data GpuOrder = RefOrder | FastOrder deriving (Eq, Show)

orderCode :: GpuOrder -> Int32
orderCode RefOrder = 0
orderCode FastOrder = 1

check :: String -> CInt -> IO ()
check what rc
  | rc == 0 = return ()
  | otherwise = do
      msg <- c_dca_last_error nullPtr >>= peekCString
      ioError (userError (what ++ ": libdca_b200 error " ++ show rc ++ ": " ++ msg))

gridBytes :: Grid -> [Word8]
gridBytes = map (\b -> if b then 1 else 0) . A.toList

frepWords :: Frep -> [Int64]
frepWords = map fromIntegral . A.toList

shape3 :: Int -> DIM3
shape3 k = Z :. rOWS :. cOLS :. k

-- | accFn1 = getAction (src/Simulator.hs:154-172)
gpuGetAction :: GpuOrder -> Scalar Cell -> Scalar Bool -> Agent -> Grid -> Frep -> (Scalar (Maybe Ch), Frep)
gpuGetAction order cellS eIsEndS (wNet, _, _) grid frep = unsafePerformIO $
  withArray (map realToFrac (A.toList wNet)) $ \pW ->
  withArray (gridBytes grid) $ \pG ->
  withArray (frepWords frep) $ \pF ->
  alloca $ \pCh ->
  allocaArray nF $ \pNext -> do
    let (r, c) = A.indexArray cellS Z
        isEnd = A.indexArray eIsEndS Z
    check "dca_acc_get_action" =<<
      c_dca_acc_get_action 0 (orderCode order) (fromIntegral r) (fromIntegral c) (if isEnd then 1 else 0)
                           pW pG pF pCh pNext
    ch <- peek pCh
    next <- peekArray nF pNext
    let mbCh = if ch < 0 then Nothing else Just (fromIntegral ch)
    return (A.fromList Z [mbCh], A.fromList (shape3 (cHANNELS + 1)) (map fromIntegral next))
  where nF = rOWS * cOLS * (cHANNELS + 1)

-- | accFn2 = gridStepTrain (src/SimRunner.hs:100-118)
gpuGridStepTrain :: GpuOrder
  -> Scalar Float -> Scalar Float -> Scalar Float -> Scalar Bool -> Scalar Cell -> Scalar (Maybe Ch)
  -> Frep -> Frep -> Grid -> Agent -> (Grid, Agent, Scalar (Maybe Float), Scalar Int)
gpuGridStepTrain order aN aA aG eIsEndS cellS mbChS frep nextFrep grid (wNet, wGrad, avgR) = unsafePerformIO $
  withArray (frepWords frep) $ \pF ->
  withArray (frepWords nextFrep) $ \pNF ->
  withArray (gridBytes grid) $ \pG ->
  withArray (map realToFrac (A.toList wNet)) $ \pW ->
  withArray (map realToFrac (A.toList wGrad)) $ \pWG ->
  alloca $ \pAvg -> alloca $ \pLoss -> alloca $ \pNaN -> alloca $ \pRew -> do
    let s x = realToFrac (A.indexArray x Z) :: CFloat
        (r, c) = A.indexArray cellS Z
        ch = maybe (-1) fromIntegral (A.indexArray mbChS Z)
    poke pAvg (s avgR)
    check "dca_acc_grid_step_train" =<<
      c_dca_acc_grid_step_train 0 (orderCode order) (s aN) (s aA) (s aG)
                                (if A.indexArray eIsEndS Z then 1 else 0) (fromIntegral r) (fromIntegral c) ch
                                pF pNF pG pW pWG pAvg pLoss pNaN pRew
    g' <- peekArray (rOWS * cOLS * cHANNELS) pG
    w' <- peekArray nF pW
    wg' <- peekArray nF pWG
    avg' <- peek pAvg
    loss <- peek pLoss
    isNaN' <- peek pNaN
    rew <- peek pRew
    let vec xs = A.fromList (Z :. nF) (map realToFrac xs) :: Array DIM1 Float
        mbLoss = if isNaN' /= 0 then Nothing else Just (realToFrac loss)  -- src/Agent.hs:80
    return ( A.fromList (shape3 cHANNELS) (map (/= 0) g')
           , (vec w', vec wg', A.fromList Z [realToFrac avg'])
           , A.fromList Z [mbLoss]
           , A.fromList Z [fromIntegral rew] )
  where nF = rOWS * cOLS * (cHANNELS + 1)

-- | featureRep (src/Gridfuncs.hs:153-170), used by mkSimState (src/Simulator.hs:96)
gpuFeatureRep :: Grid -> Frep
gpuFeatureRep grid = unsafePerformIO $
  withArray (gridBytes grid) $ \pG ->
  allocaArray nF $ \pF -> do
    check "dca_gf_feature_rep" =<< c_dca_gf_feature_rep 0 pG pF
    A.fromList (shape3 (cHANNELS + 1)) . map fromIntegral <$> peekArray nF pF
  where nF = rOWS * cOLS * (cHANNELS + 1)
