-- | The fused device path: one `dca_run` call replaces a whole period of `runStep`s
-- (runPeriod, src/SimRunner.hs:132-178, including environmentStep, src/Simulator.hs:101-134:
-- the event queue, the random numbers and the statistics live on the device).
-- `runSimGpu` is what `runSim` (src/SimRunner.hs:206-243) becomes for `--backend gpu`
-- when more than one replica is requested or `--fused` is given; reports keep the
-- reference's format because they are produced by the reference's own Stats module
-- from the counters read back.
--
-- NOT COMPILED IN THIS REPOSITORY (no GHC in the build image); see INTEGRATION.md.
module DcaB200.Fused (runSimGpu, SimStopCode(..)) where

import Control.Monad (unless, when)
import Control.Monad.Reader (MonadIO, MonadReader, ask, liftIO)
import Data.Int
import DcaB200.FFI
import Foreign.C.String (peekCString)
import Foreign.C.Types
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal.Alloc (alloca, allocaBytes)
import Foreign.Marshal.Array (allocaArray, peekArray)
import Foreign.Ptr
import Foreign.Storable (peek, pokeByteOff)
import Opt
import Stats
import Text.Printf (printf)

-- | per-replica status word = SimStop (src/SimRunner.hs:45-50)
data SimStopCode = Running | Success | ZeroLoss | NaNLoss | ReuseViolated | NEligOverflow | TapeExhausted
  deriving (Eq, Show, Enum)

-- byte offsets of dca_config (include/dca_b200.h); checked by tests/test_host_cpu.py against ctypes
offNReplicas, offSeed, offRngMode, offOrder, offNEvents, offMinLoss, offVerifyRC, offVerifyNB :: Int
offNReplicas = 0; offSeed = 8; offRngMode = 24; offOrder = 28; offNEvents = 32; offMinLoss = 40
offVerifyRC = 44; offVerifyNB = 48
offCallDur, offCallDurHoff, offCallRate, offHoffProb, offAlphaNet, offAlphaAvg, offAlphaGrad, sizeofConfig :: Int
offCallDur = 64; offCallDurHoff = 72; offCallRate = 80; offHoffProb = 88
offAlphaNet = 96; offAlphaAvg = 100; offAlphaGrad = 104; sizeofConfig = 168

withHandle :: Opt -> Int -> (Ptr DcaHandle -> IO a) -> IO a
withHandle o nReplicas act =
  allocaBytes sizeofConfig $ \cfg -> alloca $ \pH -> do
    _ <- c_dca_config_defaults cfg
    pokeByteOff cfg offNReplicas (fromIntegral nReplicas :: Int32)
    pokeByteOff cfg offSeed (rngSeed o)
    -- --fixed_rng replays the reference's MT19937-64 stream; otherwise Philox keyed (seed, replica)
    pokeByteOff cfg offRngMode (if nReplicas == 1 then 0 else 1 :: Int32)
    pokeByteOff cfg offNEvents (fromIntegral (nEvents o) :: Int64)
    pokeByteOff cfg offMinLoss (realToFrac (minLoss o) :: CFloat)
    pokeByteOff cfg offVerifyRC (if verifyReuseConstraint o then 1 else 0 :: Int32)
    pokeByteOff cfg offVerifyNB (if verifyNEligBounds o then 1 else 0 :: Int32)
    pokeByteOff cfg offCallDur (callDurNew o)
    pokeByteOff cfg offCallDurHoff (callDurHoff o)
    pokeByteOff cfg offCallRate (callRate o)
    pokeByteOff cfg offHoffProb (hoffProb o)
    pokeByteOff cfg offAlphaNet (realToFrac (alphaNet o) :: CFloat)
    pokeByteOff cfg offAlphaAvg (realToFrac (alphaAvg o) :: CFloat)
    pokeByteOff cfg offAlphaGrad (realToFrac (alphaGrad o) :: CFloat)
    rc <- c_dca_create cfg pH
    when (rc /= 0) $ c_dca_last_error nullPtr >>= peekCString >>= ioError . userError
    h <- peek pH
    fp <- newForeignPtr (castFunPtr p_dca_destroy) h :: IO (ForeignPtr DcaHandle)
    when (nReplicas == 1) $ do
      _ <- c_dca_set_rng_mt19937 h 0 (rngSeed o) (fromIntegral (4 * nEvents o + 64))
      return ()
    withForeignPtr fp act

-- | `runSim` for the device path: periods of `logIter` events, the reference's reports.
runSimGpu :: (MonadReader Opt m, MonadIO m) => Int -> m ()
runSimGpu nReplicas = do
  o <- ask
  liftIO $ withHandle o nReplicas $ \h -> do
    let loop = do
          _ <- c_dca_run h (fromIntegral (logIter o))  -- one period, src/SimRunner.hs:167-168
          (status, iter) <- alloca $ \pS -> alloca $ \pI -> do
            _ <- c_dca_read_status h 0 pS pI nullPtr
            (,) <$> peek pS <*> peek pI
          counters <- allocaArray 8 $ \p -> c_dca_read_stats h 0 p >> peekArray 8 p
          (lossSum, lossN) <- alloca $ \pL -> alloca $ \pN -> do
            _ <- c_dca_read_period h 0 pL pN 1
            (,) <$> peek pL <*> peek pN
          avgR <- alloca $ \pA -> do
            _ <- c_dca_read_state h 0 nullPtr nullPtr nullPtr nullPtr pA nullPtr nullPtr
            peek pA
          -- statsReportPeriod's line (src/Stats.hs:94-124) from the device counters
          putStrLn (periodReport counters iter (logIter o) (realToFrac (lossSum :: CFloat)) (lossN :: Int64)
                                 (realToFrac (avgR :: CFloat)))
          let stop = toEnum (fromIntegral (status :: Int32)) :: SimStopCode
          unless (stop /= Running) loop
          when (stop /= Running) $ print stop
    loop
    total <- allocaArray 10 $ \p -> c_dca_sum_stats h p nullPtr >> peekArray 10 p
    ms <- alloca $ \p -> c_dca_elapsed_ms h p >> peek p
    printf "%d events on the device; last period %.1f ms\n" (total !! 8 :: Int64) (realToFrac (ms :: CFloat) :: Double)

-- | `Stats.hs:108-115` with the counters laid out as in include/dca_b200.h (DCA_ST_*)
periodReport :: [Int64] -> Int64 -> Int -> Double -> Int64 -> Double -> String
periodReport st iter lgIter lossSum lossN avgReward =
  printf "Blocking probability events %d-%d: %.4f, cumulative %.4f, avg. loss: %.1f, avg. reward %.4f"
         (max (iter - fromIntegral lgIter) 0) iter
         (fromIntegral (st !! 7) / (fromIntegral (st !! 6) + 1) :: Double)
         (fromIntegral (st !! 2) / (fromIntegral (st !! 0) + 1) :: Double)
         (lossSum / fromIntegral lossN :: Double)
         avgReward
