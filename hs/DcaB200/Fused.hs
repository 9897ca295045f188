-- | The fused device path: one `dca_run` call replaces a whole period of `runStep`s
-- (runPeriod, src/SimRunner.hs:132-178, including environmentStep, src/Simulator.hs:101-134:
-- the event queue, the random numbers and the statistics live on the device).
-- `runSimGpu` is what `runSim` (src/SimRunner.hs:206-243) becomes for `--backend gpu`
-- when more than one replica is requested or `--fused` is given; the period report keeps the
-- reference's format (`periodReport` below restates `Stats.hs:108-115` on the counters read back).
--
-- EXPERIMENTAL / NOT COMPILED IN THIS REPOSITORY (no GHC in the build image); see INTEGRATION.md.  Until it
-- has been built with GHC 8.4 treat it as a reviewed sketch: the C ABI underneath is what is tested.
module DcaB200.Fused (runSimGpu, RngSource(..), SimStopCode(..)) where

import Control.Monad (forM_, when)
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
import Text.Printf (printf)

-- | per-replica status word = SimStop (src/SimRunner.hs:45-50)
data SimStopCode = Running | Success | ZeroLoss | NaNLoss | ReuseViolated | NEligOverflow | TapeExhausted
                 | Uninitialized | EventQueueOverflow   -- DCA_UNINITIALIZED, DCA_QUEUE_OVERFLOW (include/dca_b200.h)
  deriving (Eq, Show, Enum)

-- byte offsets of dca_config (include/dca_b200.h); checked by tests/test_host_cpu.py against ctypes
offNReplicas, offSeed, offRngMode, offOrder, offNEvents, offMinLoss, offVerifyRC, offVerifyNB :: Int
offNReplicas = 0; offSeed = 8; offRngMode = 24; offOrder = 28; offNEvents = 32; offMinLoss = 40
offVerifyRC = 44; offVerifyNB = 48
offCallDur, offCallDurHoff, offCallRate, offHoffProb, offAlphaNet, offAlphaAvg, offAlphaGrad, sizeofConfig :: Int
offCallDur = 64; offCallDurHoff = 72; offCallRate = 80; offHoffProb = 88
offAlphaNet = 96; offAlphaAvg = 100; offAlphaGrad = 104; sizeofConfig = 176

-- | every FFI return code goes through here: a failed launch (not an sm_100a device, tensor-memory allocation
-- failure, bad handle, a replica without its random tape ...) aborts instead of leaving a status at Running
check :: Ptr DcaHandle -> String -> CInt -> IO ()
check h what rc
  | rc == 0 = return ()
  | otherwise = do
      msg <- c_dca_last_error h >>= peekCString
      ioError (userError (what ++ ": libdca_b200 error " ++ show rc ++ ": " ++ msg))

-- | which random source the device uses: the reference's MT19937-64 stream replayed from a host tape
-- (`--fixed_rng`, src/Opt.hs:79-82; one tape per replica, seeds rngSeed + r) or Philox keyed (seed, replica)
data RngSource = FixedRngTape | PhiloxStreams deriving (Eq, Show)

withHandle :: Opt -> RngSource -> Int -> (Ptr DcaHandle -> IO a) -> IO a
withHandle o rng nReplicas act =
  allocaBytes sizeofConfig $ \cfg -> alloca $ \pH -> do
    check nullPtr "dca_config_defaults" =<< c_dca_config_defaults cfg
    pokeByteOff cfg offNReplicas (fromIntegral nReplicas :: Int32)
    pokeByteOff cfg offSeed (rngSeed o)
    pokeByteOff cfg offRngMode (if rng == FixedRngTape then 0 else 1 :: Int32)
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
    check nullPtr "dca_create" =<< c_dca_create cfg pH
    h <- peek pH
    fp <- newForeignPtr (castFunPtr p_dca_destroy) h :: IO (ForeignPtr DcaHandle)
    withForeignPtr fp $ \h' -> do
      when (rng == FixedRngTape) $
        forM_ [0 .. nReplicas - 1] $ \r ->
          check h' "dca_set_rng_mt19937" =<<
            c_dca_set_rng_mt19937 h' (fromIntegral r) (rngSeed o + fromIntegral r) (fromIntegral (4 * nEvents o + 64))
      act h'

-- | statuses of all replicas (dca_read_summary); the run is over when none is Running
replicaStatuses :: Ptr DcaHandle -> Int -> IO [SimStopCode]
replicaStatuses h n = allocaArray n $ \pS -> do
  check h "dca_read_summary" =<< c_dca_read_summary h pS nullPtr nullPtr nullPtr
  map (toEnum . fromIntegral) <$> (peekArray n pS :: IO [Int32])

-- | `runSim` for the device path: periods of `logIter` events, the reference's reports for replica 0.
-- The loop ends when EVERY replica has left Running (a replica that stops early -- ZeroLoss, NaNLoss -- does
-- not end the others), and aborts if a period made no progress at all.
runSimGpu :: (MonadReader Opt m, MonadIO m) => RngSource -> Int -> m ()
runSimGpu rng nReplicas = do
  o <- ask
  liftIO $ withHandle o rng nReplicas $ \h -> do
    let loop lastTotal = do
          check h "dca_run" =<< c_dca_run h (fromIntegral (logIter o))  -- one period, src/SimRunner.hs:167-168
          (status, iter) <- alloca $ \pS -> alloca $ \pI -> do
            check h "dca_read_status" =<< c_dca_read_status h 0 pS pI nullPtr
            (,) <$> peek pS <*> peek pI
          counters <- allocaArray 8 $ \p -> (check h "dca_read_stats" =<< c_dca_read_stats h 0 p) >> peekArray 8 p
          (lossSum, lossN) <- alloca $ \pL -> alloca $ \pN -> do
            check h "dca_read_period" =<< c_dca_read_period h 0 pL pN 1
            (,) <$> peek pL <*> peek pN
          avgR <- alloca $ \pA -> do
            check h "dca_read_state" =<< c_dca_read_state h 0 nullPtr nullPtr nullPtr nullPtr pA nullPtr nullPtr
            peek pA
          let stop0 = toEnum (fromIntegral (status :: Int32)) :: SimStopCode
          -- statsReportPeriod's line (src/Stats.hs:94-124) from the device counters, while replica 0 still moves
          when (lossN > 0) $
            putStrLn (periodReport counters iter (logIter o) (realToFrac (lossSum :: CFloat)) (lossN :: Int64)
                                   (realToFrac (avgR :: CFloat)))
          when (stop0 /= Running && lossN > 0) $ print stop0
          statuses <- replicaStatuses h nReplicas
          total <- allocaArray 10 $ \p -> (check h "dca_sum_stats" =<< c_dca_sum_stats h p nullPtr) >> peekArray 10 p
          let processed = total !! 8 :: Int64
          when (any (== Running) statuses) $ do
            when (processed == lastTotal) $ ioError (userError "runSimGpu: a period made no progress")
            loop processed
    loop (-1)
    total <- allocaArray 10 $ \p -> (check h "dca_sum_stats" =<< c_dca_sum_stats h p nullPtr) >> peekArray 10 p
    ms <- alloca $ \p -> (check h "dca_elapsed_ms" =<< c_dca_elapsed_ms h p) >> peek p
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
