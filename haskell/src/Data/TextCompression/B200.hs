{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE MultiWayIf               #-}

-- |
-- Module      : Data.TextCompression.B200
-- Description : `foreign import ccall` shim over libtc_b200.so (include/tc_b200.h)
--
-- The thin host layer the text-compression package binds to reach the B200 kernels.
-- Every function here replaces one function of the reference's "Internal" modules and keeps
-- its type, so the public modules (Data.BWT, Data.MTF, Data.RLE, Data.FMIndex) and their
-- export lists stay as they are (see INTEGRATION.md for the per-function patch).
--
-- NOTE: written without a Haskell toolchain (no GHC in the build image or on the GPU box);
-- it targets base ^>=4.12 / containers 0.6 like the reference, but HAS NOT BEEN COMPILED.  What can be checked
-- without GHC is checked: tests/c/abi_layout.c restates every foreign import below as a C prototype against
-- include/tc_b200.h and pins the struct offsets used here (544 / 552 / 16) with offsetof / sizeof.
-- All semantics live below the C ABI, where they are tested bit-for-bit against the CPU
-- restatement of the reference (tests/test_gpu_parity.py).
--
-- Marshalling: `Seq (Maybe b)` over byte-derived symbols <-> int16 (-1 == Nothing);
-- `Seq Int` <-> uint16 / int64; buffers are pinned (tc_host_alloc) so the H2D/D2H copies inside
-- the library run at full PCIe rate.  Error codes map back to the exceptions the reference
-- throws: TC_E_FROMJUST -> `fromJust Nothing`, TC_E_INDEX -> `index out of bounds`.
module Data.TextCompression.B200
  ( -- * context
    withB200
  , deviceCount
    -- * Data.BWT.Internal replacements
  , createSuffixArrayW8
  , toBWTW8
  , fromBWTW8
    -- * Data.MTF.Internal replacements
  , seqToMTFW8
  , seqFromMTFW8
    -- * Data.RLE.Internal replacements
  , seqToRLEW8
  , seqFromRLEW8
    -- * composed helpers (device-resident BWT -> MTF -> RLE)
  , bwtMtfRleW8
    -- * multi-block compression with packed block containers (tc_packed_header, include/tc_b200.h)
  , compressBlocksPackedW8
  , unpackBlockW8
  , decodePackedW8
  , decodeBlocksPackedW8
    -- * unboxed results alongside the Seq API (SURVEY.md 8f.1): packed arrays in strict ByteStrings, no per-element boxing
  , PackedBWT (..)
  , toBWTPackedW8
  , fromBWTPackedW8
  , toMTFPackedW8
    -- * Data.FMIndex.Internal replacements
  , B200FM
  , buildFMIndexW8
  , countFMIndexW8
  , locateFMIndexW8
    -- * the same over every GPU of the box (one library call; mirrors parListChunk over the cores)
  , compressBlocksPackedMultiW8
  , B200FMs
  , replicateFMIndexW8
  , countFMIndexMultiW8
  , locateFMIndexMultiW8
  ) where

import           Control.Exception     (ErrorCall (..), bracket, throwIO)
import           Control.Monad         (forM, forM_, when)
import qualified Data.ByteString       as BS
import qualified Data.ByteString.Unsafe as BSU
import           Data.Foldable         (toList)
import           Data.Int              (Int16, Int64)
import           Data.Sequence         (Seq)
import qualified Data.Sequence         as DS
import           Data.Word             (Word16, Word32, Word64, Word8)
import           Foreign.C.String      (CString, peekCString)
import           Foreign.C.Types       (CInt (..), CSize (..))
import           Foreign.ForeignPtr    (ForeignPtr, newForeignPtr, touchForeignPtr, withForeignPtr)
import           Foreign.Marshal.Alloc (alloca)
import           Foreign.Marshal.Array (allocaArray, peekArray, pokeArray, withArray)
import           Foreign.Marshal.Utils (copyBytes, withMany)
import           Foreign.Ptr           (FunPtr, Ptr, castPtr, nullPtr)
import           Foreign.Storable      (peek, peekByteOff, peekElemOff, pokeElemOff)
import           System.Environment   (lookupEnv)
import           System.IO.Unsafe      (unsafePerformIO)

data TcCtx
data TcFm

-- include/tc_b200.h ------------------------------------------------------------------
foreign import ccall safe "tc_ctx_pool_acquire" c_pool_acquire :: CInt -> Ptr (Ptr TcCtx) -> IO CInt
foreign import ccall safe "tc_ctx_pool_release" c_pool_release :: Ptr TcCtx -> IO ()
foreign import ccall unsafe "tc_device_count"   c_device_count :: IO CInt
foreign import ccall safe "tc_mgpu_blocks_encode_packed"
  c_mgpu_blocks_encode_packed :: CInt -> Ptr CInt -> Word64 -> Ptr (Ptr Word8) -> Ptr Word64 -> CInt -> Ptr (Ptr Word8)
                              -> Ptr Word64 -> Ptr Word64 -> Ptr () -> IO CInt
foreign import ccall safe "tc_fm_replicate"
  c_fm_replicate :: Ptr TcFm -> CInt -> Ptr CInt -> Ptr (Ptr TcFm) -> IO CInt
foreign import ccall safe "tc_mgpu_fm_count"
  c_mgpu_fm_count :: CInt -> Ptr CInt -> Ptr (Ptr TcFm) -> Ptr Word8 -> Ptr Word64 -> Word64 -> Ptr Int64 -> IO CInt
foreign import ccall safe "tc_mgpu_fm_locate"
  c_mgpu_fm_locate :: CInt -> Ptr CInt -> Ptr (Ptr TcFm) -> Ptr Word8 -> Ptr Word64 -> Word64 -> Ptr Word64
                   -> Ptr Word64 -> Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_strerror"     c_strerror    :: CInt -> IO CString
foreign import ccall safe "tc_last_error"   c_last_error  :: Ptr TcCtx -> IO CString
foreign import ccall safe "tc_host_alloc"   c_host_alloc  :: CSize -> IO (Ptr a)
foreign import ccall safe "&tc_host_free"   p_host_free   :: FunPtr (Ptr a -> IO ())
foreign import ccall safe "tc_bwt_encode"
  c_bwt_encode :: Ptr TcCtx -> Ptr Word8 -> Word64 -> Ptr Word8 -> Ptr Word64 -> Ptr Word32 -> IO CInt
foreign import ccall safe "tc_bwt_decode"
  c_bwt_decode :: Ptr TcCtx -> Ptr Int16 -> Word64 -> Ptr Word8 -> Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_mtf_encode"
  c_mtf_encode :: Ptr TcCtx -> Ptr Int16 -> Word64 -> Ptr Word16 -> Ptr Int16 -> Ptr Word32 -> IO CInt
foreign import ccall safe "tc_bwt_decode_u8"
  c_bwt_decode_u8 :: Ptr TcCtx -> Ptr Word8 -> Word64 -> Word64 -> Ptr Word8 -> Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_mtf_encode_u8"
  c_mtf_encode_u8 :: Ptr TcCtx -> Ptr Word8 -> Word64 -> Word64 -> Ptr Word16 -> Ptr Int16 -> Ptr Word32 -> IO CInt
foreign import ccall safe "tc_mtf_decode"
  c_mtf_decode :: Ptr TcCtx -> Ptr Word16 -> Word64 -> Ptr Int16 -> Word32 -> Ptr Int16 -> IO CInt
foreign import ccall safe "tc_rle_encode"
  c_rle_encode :: Ptr TcCtx -> Ptr Int16 -> Word64 -> Ptr Word32 -> Ptr Int16 -> Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_rle_decode"
  c_rle_decode :: Ptr TcCtx -> Ptr Word32 -> Ptr Int16 -> Word64 -> Ptr Int16 -> Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_bwt_mtf_rle_encode"
  c_bwt_mtf_rle :: Ptr TcCtx -> Ptr Word8 -> Word64 -> Ptr Word32 -> Ptr Int16 -> Word64 -> Ptr () -> IO CInt
-- multi-block entry point: copies of neighbouring blocks overlap the kernels of the current one
foreign import ccall safe "tc_blocks_encode"
  c_blocks_encode :: Ptr TcCtx -> Word64 -> Ptr (Ptr Word8) -> Ptr Word64 -> CInt -> Ptr (Ptr Word32)
                  -> Ptr (Ptr Int16) -> Ptr Word64 -> Ptr () -> IO CInt
-- same, with one packed block container per block as output (tc_packed_header + cnt4 / sym8 / hi
-- sections): a third of the bytes over PCIe, and a single ByteString per block on the Haskell side
foreign import ccall safe "tc_packed_bound" c_packed_bound :: Word64 -> IO Word64
foreign import ccall safe "tc_blocks_encode_packed"
  c_blocks_encode_packed :: Ptr TcCtx -> Word64 -> Ptr (Ptr Word8) -> Ptr Word64 -> CInt -> Ptr (Ptr Word8)
                         -> Ptr Word64 -> Ptr Word64 -> Ptr () -> IO CInt
-- host only: container -> the (count, symbol) records of tc_blocks_encode
foreign import ccall unsafe "tc_packed_unpack"
  c_packed_unpack :: Ptr Word8 -> Word64 -> Ptr Word32 -> Ptr Int16 -> Word64 -> Ptr () -> IO CInt
foreign import ccall safe "tc_packed_decode"
  c_packed_decode :: Ptr TcCtx -> Ptr Word8 -> Word64 -> Ptr Word8 -> Word64 -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_blocks_decode_packed"
  c_blocks_decode_packed :: Ptr TcCtx -> Word64 -> Ptr (Ptr Word8) -> Ptr Word64 -> Ptr (Ptr Word8) -> Ptr Word64
                         -> Ptr Word64 -> IO CInt
foreign import ccall safe "tc_fm_build"
  c_fm_build :: Ptr TcCtx -> Ptr Word8 -> Word64 -> Word32 -> Ptr (Ptr TcFm) -> IO CInt
foreign import ccall safe "&tc_fm_free" p_fm_free :: FunPtr (Ptr TcFm -> IO ())
foreign import ccall safe "tc_fm_count"
  c_fm_count :: Ptr TcCtx -> Ptr TcFm -> Ptr Word8 -> Ptr Word64 -> Word64 -> Ptr Int64 -> IO CInt
foreign import ccall safe "tc_fm_locate"
  c_fm_locate :: Ptr TcCtx -> Ptr TcFm -> Ptr Word8 -> Ptr Word64 -> Word64 -> Ptr Word64 -> Ptr Word64
              -> Word64 -> Ptr Word64 -> IO CInt

tcECap, tcEFromJust, tcEIndex :: CInt
tcECap      = -2
tcEFromJust = -3
tcEIndex    = -4

-- | The process-wide context of the device (tc_ctx_pool_acquire: created on first use, exclusive while the
-- action runs, never destroyed), so a pure call costs no stream / arena / pinned-page set-up.  The device is
-- TC_B200_DEVICE (default 0).  Calls from several Haskell threads serialise on the device's context; the
-- ...Multi functions below use one context per GPU.
withB200 :: (Ptr TcCtx -> IO a) -> IO a
withB200 = bracket acquire c_pool_release
  where
    acquire = alloca $ \pp -> do
      dev <- maybe 0 read <$> lookupEnv "TC_B200_DEVICE"
      rc  <- c_pool_acquire dev pp
      when (rc /= 0) $ c_strerror rc >>= peekCString >>= throwIO . ErrorCall
      peek pp

-- | GPUs visible to the process (0: none, and every compute call throws -- there is no CPU path).
deviceCount :: Int
deviceCount = fromIntegral (unsafePerformIO c_device_count)

checkRc :: CInt -> IO ()
checkRc rc
  | rc == 0           = pure ()
  | rc == tcEFromJust = throwIO (ErrorCall "Maybe.fromJust: Nothing")
  | rc == tcEIndex    = throwIO (ErrorCall "index out of bounds")
  | otherwise         = c_strerror rc >>= peekCString >>= throwIO . ErrorCall . ("libtc_b200: " ++)

check :: Ptr TcCtx -> CInt -> IO ()
check ctx rc
  | rc == 0           = pure ()
  | rc == tcEFromJust = throwIO (ErrorCall "Maybe.fromJust: Nothing")
  | rc == tcEIndex    = throwIO (ErrorCall "index out of bounds")
  | otherwise         = do
      s <- c_strerror rc >>= peekCString
      d <- c_last_error ctx >>= peekCString
      throwIO (ErrorCall ("libtc_b200: " ++ s ++ " " ++ d))

pinned :: Int -> IO (ForeignPtr a)
pinned bytes = c_host_alloc (fromIntegral (max 1 bytes)) >>= newForeignPtr p_host_free

maybeToI16 :: Maybe Word8 -> Int16
maybeToI16 Nothing  = -1
maybeToI16 (Just w) = fromIntegral w

i16ToMaybe :: Int16 -> Maybe Word8
i16ToMaybe x | x < 0     = Nothing
             | otherwise = Just (fromIntegral x)

-- | createSuffixArray (Data/BWT/Internal.hs:110-134) for Word8 text: suffixstartpos per rank.
createSuffixArrayW8 :: BS.ByteString -> Seq Int
createSuffixArrayW8 xs
  | BS.null xs = DS.singleton 1
  | otherwise  = unsafePerformIO $ withB200 $ \ctx ->
      BSU.unsafeUseAsCStringLen xs $ \(p, n) -> do
        bwt <- pinned (n + 1)
        sa  <- pinned (4 * (n + 1))
        alloca $ \pprim -> withForeignPtr bwt $ \pb -> withForeignPtr sa $ \ps -> do
          c_bwt_encode ctx (castPtr p) (fromIntegral n) pb pprim ps >>= check ctx
          DS.fromList . map fromIntegral <$> (peekArray (n + 1) ps :: IO [Word32])

-- | toBWT (Data/BWT.hs:55-64) on a ByteString: the BWT as `Seq (Maybe Word8)`.
toBWTW8 :: BS.ByteString -> Seq (Maybe Word8)
toBWTW8 xs
  | BS.null xs = DS.empty
  | otherwise  = unsafePerformIO $ withB200 $ \ctx ->
      BSU.unsafeUseAsCStringLen xs $ \(p, n) -> do
        bwt <- pinned (n + 1)
        alloca $ \pprim -> withForeignPtr bwt $ \pb -> do
          c_bwt_encode ctx (castPtr p) (fromIntegral n) pb pprim nullPtr >>= check ctx
          prim <- fromIntegral <$> peek pprim
          ws   <- peekArray (n + 1) pb
          pure $ DS.fromList [ if i == prim then Nothing else Just w | (i, w) <- zip [0 :: Int ..] ws ]

-- | fromBWT (Data/BWT.hs:93-104) on `Seq (Maybe Word8)`.
fromBWTW8 :: Seq (Maybe Word8) -> BS.ByteString
fromBWTW8 s
  | DS.null s = BS.empty
  | otherwise = unsafePerformIO $ withB200 $ \ctx -> do
      let n = DS.length s
      inp <- pinned (2 * n)
      out <- pinned n
      alloca $ \pn -> withForeignPtr inp $ \pi' -> withForeignPtr out $ \po -> do
        pokeArray pi' (map maybeToI16 (toList s))
        c_bwt_decode ctx pi' (fromIntegral n) po (fromIntegral n) pn >>= check ctx
        m <- fromIntegral <$> peek pn
        BS.packCStringLen (castPtr po, m)

-- | seqToMTF (Data/MTF/Internal.hs:128-175): (indices, FINAL list).
seqToMTFW8 :: Seq (Maybe Word8) -> (Seq Int, Seq (Maybe Word8))
seqToMTFW8 s
  | DS.null s = (DS.empty, DS.empty)
  | otherwise = unsafePerformIO $ withB200 $ \ctx -> do
      let n = DS.length s
      inp <- pinned (2 * n)
      idx <- pinned (2 * n)
      fin <- pinned (2 * 257)
      alloca $ \psig -> withForeignPtr inp $ \pi' -> withForeignPtr idx $ \px -> withForeignPtr fin $ \pf -> do
        pokeArray pi' (map maybeToI16 (toList s))
        c_mtf_encode ctx pi' (fromIntegral n) px pf psig >>= check ctx
        sg <- fromIntegral <$> peek psig
        is <- peekArray n px :: IO [Word16]
        fs <- peekArray sg pf
        pure (DS.fromList (map fromIntegral is), DS.fromList (map i16ToMaybe fs))

-- | seqFromMTF (Data/MTF/Internal.hs:201-232).
seqFromMTFW8 :: (Seq Int, Seq (Maybe Word8)) -> Seq (Maybe Word8)
seqFromMTFW8 (is, fl)
  | DS.null is || DS.null fl = DS.empty
  | otherwise = unsafePerformIO $ withB200 $ \ctx -> do
      let n = DS.length is
          sg = DS.length fl
      idx <- pinned (2 * n)
      fin <- pinned (2 * sg)
      out <- pinned (2 * n)
      withForeignPtr idx $ \px -> withForeignPtr fin $ \pf -> withForeignPtr out $ \po -> do
        forM_ (toList is) $ \i -> when (i < 0 || i > 65535) $ throwIO (ErrorCall "index out of bounds")
        pokeArray px (map fromIntegral (toList is) :: [Word16])
        pokeArray pf (map maybeToI16 (toList fl))
        c_mtf_decode ctx px (fromIntegral n) pf (fromIntegral sg) po >>= check ctx
        DS.fromList . map i16ToMaybe <$> peekArray n po

-- | seqToRLE (Data/RLE/Internal.hs:104-153) as (count, symbol) runs; the flat
-- `Seq (Maybe b)` of the reference is @[Just (fromString (show c)), sym]@ per run.
seqToRLEW8 :: Seq (Maybe Word8) -> Seq (Int, Maybe Word8)
seqToRLEW8 s
  | DS.null s = DS.empty
  | otherwise = unsafePerformIO $ withB200 $ \ctx -> do
      let n   = DS.length s
          cap = 2 * n + 1
      inp <- pinned (2 * n)
      cnt <- pinned (4 * cap)
      sym <- pinned (2 * cap)
      alloca $ \pr -> withForeignPtr inp $ \pi' -> withForeignPtr cnt $ \pc -> withForeignPtr sym $ \ps -> do
        pokeArray pi' (map maybeToI16 (toList s))
        c_rle_encode ctx pi' (fromIntegral n) pc ps (fromIntegral cap) pr >>= check ctx
        r  <- fromIntegral <$> peek pr
        cs <- peekArray r pc :: IO [Word32]
        ss <- peekArray r ps
        pure (DS.fromList (zip (map fromIntegral cs) (map i16ToMaybe ss)))

-- | seqFromRLE (Data/RLE/Internal.hs:155-189) from (count, symbol) runs.
seqFromRLEW8 :: Seq (Int, Maybe Word8) -> Seq (Maybe Word8)
seqFromRLEW8 rs
  | DS.null rs = DS.empty
  | otherwise  = unsafePerformIO $ withB200 $ \ctx -> do
      let r = DS.length rs
      cnt <- pinned (4 * r)
      sym <- pinned (2 * r)
      alloca $ \pn -> withForeignPtr cnt $ \pc -> withForeignPtr sym $ \ps -> do
        pokeArray pc (map (fromIntegral . max 0 . fst) (toList rs) :: [Word32])
        pokeArray ps (map (maybeToI16 . snd) (toList rs))
        rc <- c_rle_decode ctx pc ps (fromIntegral r) nullPtr 0 pn      -- sizing call
        when (rc /= 0 && rc /= tcECap) $ check ctx rc
        n <- fromIntegral <$> peek pn
        out <- pinned (2 * n)
        withForeignPtr out $ \po -> do
          c_rle_decode ctx pc ps (fromIntegral r) po (fromIntegral n) pn >>= check ctx
          DS.fromList . map i16ToMaybe <$> peekArray n po

-- | bytestringToBWTToMTFB followed by seqToRLE over the index stream, chained in HBM
-- (SURVEY.md 8b).  Returns the runs (count, MTF index); tc_block_info carries primary,
-- sigma and the MTF final list (read it with hsc2hs/offsets in a full binding).
bwtMtfRleW8 :: BS.ByteString -> Seq (Int, Int)
bwtMtfRleW8 xs
  | BS.null xs = DS.empty
  | otherwise  = unsafePerformIO $ withB200 $ \ctx ->
      BSU.unsafeUseAsCStringLen xs $ \(p, n) -> do
        let cap = n + 3
        cnt  <- pinned (4 * cap)
        sym  <- pinned (2 * cap)
        info <- pinned 1024
        withForeignPtr cnt $ \pc -> withForeignPtr sym $ \ps -> withForeignPtr info $ \pinfo -> do
          c_bwt_mtf_rle ctx (castPtr p) (fromIntegral n) pc ps (fromIntegral cap) pinfo >>= check ctx
          -- tc_block_info: n, N, primary (3 x u64), sigma (u32), final_list (257 x i16) = 542 bytes,
          -- padded to 544; R follows (include/tc_b200.h).  A full binding would use hsc2hs.
          r  <- fromIntegral <$> (peekByteOff pinfo 544 :: IO Word64)
          cs <- peekArray r pc :: IO [Word32]
          ss <- peekArray r ps :: IO [Int16]
          pure (DS.fromList (zip (map fromIntegral cs) (map fromIntegral ss)))

-- | @fmap bytestringToBWTToMTFB blocks@ + RLE of the index streams, kept compressed: one packed block
-- container (a strict ByteString: tc_packed_header, then the runs at 2 bytes + 1 bit each) per
-- input block, in input order.  One call for the whole list: the library overlaps every block's
-- H2D / D2H copies with the kernels of its neighbours and keeps several blocks in flight.
compressBlocksPackedW8 :: [BS.ByteString] -> [BS.ByteString]
compressBlocksPackedW8 []     = []
compressBlocksPackedW8 blocks = unsafePerformIO $ withB200 $ \ctx -> do
  let nb = length blocks
      ns = map (fromIntegral . BS.length) blocks :: [Word64]
  caps <- mapM c_packed_bound ns
  outs <- mapM (pinned . fromIntegral) caps :: IO [ForeignPtr Word8]
  info <- pinned (552 * nb)                                  -- tc_block_info is 552 bytes
  withMany BSU.unsafeUseAsCStringLen blocks $ \ins ->
    withMany withForeignPtr outs $ \pouts ->
      withArray (map (castPtr . fst) ins) $ \ptext ->
        withArray ns $ \pn -> withArray pouts $ \pout -> withArray caps $ \pcap ->
          withArray (replicate nb 0) $ \pbytes -> withForeignPtr info $ \pinfo -> do
            c_blocks_encode_packed ctx (fromIntegral nb) ptext pn 1 pout pcap pbytes pinfo >>= check ctx
            sizes <- peekArray nb pbytes
            forM (zip pouts sizes) $ \(po, sz) -> BS.packCStringLen (castPtr po, fromIntegral sz)

-- | Host only (no device): the (count, MTF index or BWT symbol) runs a container holds -- what
-- `bwtMtfRleW8` returns for the same block.  A symbol of -1 is `Nothing`.
unpackBlockW8 :: BS.ByteString -> Seq (Int, Int)
unpackBlockW8 blob = unsafePerformIO $ BSU.unsafeUseAsCStringLen blob $ \(p, len) -> do
  info <- pinned 552
  withForeignPtr info $ \pinfo -> do
    rc0 <- c_packed_unpack (castPtr p) (fromIntegral len) nullPtr nullPtr 0 pinfo   -- header only: R
    when (rc0 /= 0 && rc0 /= tcECap) $ throwIO (ErrorCall "libtc_b200: malformed block container")
    r <- fromIntegral <$> (peekByteOff pinfo 544 :: IO Word64)
    if r == 0 then pure DS.empty else do
      cnt <- pinned (4 * r)
      sym <- pinned (2 * r)
      withForeignPtr cnt $ \pc -> withForeignPtr sym $ \ps -> do
        rc <- c_packed_unpack (castPtr p) (fromIntegral len) pc ps (fromIntegral r) pinfo
        when (rc /= 0) $ throwIO (ErrorCall "libtc_b200: malformed block container")
        cs <- peekArray r pc :: IO [Word32]
        ss <- peekArray r ps :: IO [Int16]
        pure (DS.fromList (zip (map fromIntegral cs) (map fromIntegral ss)))

-- | Container -> text, unpacked and decoded on the device (the inverse chain of
-- bytestringFromBWTFromMTFB / bytestringFromBWTFromRLEB, src/Data/MTF.hs, src/Data/RLE.hs).
decodePackedW8 :: BS.ByteString -> BS.ByteString
decodePackedW8 blob = unsafePerformIO $ withB200 $ \ctx ->
  BSU.unsafeUseAsCStringLen blob $ \(p, len) -> do
    n <- if len >= 24 then fromIntegral <$> (peekByteOff p 16 :: IO Word64) else pure 0   -- tc_packed_header.n
    out <- pinned (n + 2)
    alloca $ \pn -> withForeignPtr out $ \po -> do
      c_packed_decode ctx (castPtr p) (fromIntegral len) po (fromIntegral (n + 2)) pn >>= check ctx
      m <- fromIntegral <$> peek pn
      BS.packCStringLen (castPtr po, m)

-- | Multi-block decompression (tc_blocks_decode_packed): several containers in flight on the device.
decodeBlocksPackedW8 :: [BS.ByteString] -> [BS.ByteString]
decodeBlocksPackedW8 []    = []
decodeBlocksPackedW8 blobs = unsafePerformIO $ withB200 $ \ctx -> do
  let nb = length blobs
  ns <- forM blobs $ \b -> BSU.unsafeUseAsCStringLen b $ \(p, len) ->
          if len >= 24 then fromIntegral <$> (peekByteOff p 16 :: IO Word64) else pure (0 :: Int)   -- tc_packed_header.n
  outs <- forM ns $ \n -> pinned (n + 2)
  ins  <- forM blobs $ \b -> do                      -- containers in pinned memory: the H2D copies run at PCIe rate
            fp <- pinned (BS.length b)
            withForeignPtr fp $ \d -> BSU.unsafeUseAsCStringLen b $ \(p, len) -> copyBytes d (castPtr p) len
            pure fp
  allocaArray nb $ \pblob -> allocaArray nb $ \pbytes -> allocaArray nb $ \ptext ->
    allocaArray nb $ \pcap -> allocaArray nb $ \pn -> do
      forM_ (zip3 [0 ..] ins blobs) $ \(i, fp, b) -> withForeignPtr fp $ \d -> do
        pokeElemOff pblob i d
        pokeElemOff pbytes i (fromIntegral (BS.length b))
      forM_ (zip3 [0 ..] outs ns) $ \(i, fp, n) -> withForeignPtr fp $ \d -> do
        pokeElemOff ptext i d
        pokeElemOff pcap i (fromIntegral (n + 2))
      c_blocks_decode_packed ctx (fromIntegral nb) pblob pbytes ptext pcap pn >>= check ctx
      res <- forM (zip [0 ..] outs) $ \(i, fp) -> do
        m <- fromIntegral <$> peekElemOff pn i
        withForeignPtr fp $ \d -> BS.packCStringLen (castPtr d, m)
      mapM_ touchForeignPtr ins
      pure res

-- | The BWT of a ByteString without a `Seq (Maybe Word8)`: the n + 1 column bytes (the byte in slot `bwtPrimary` is
-- unspecified: that slot is the Nothing) in one strict ByteString.  8f.1: results stay packed until a caller asks
-- for the `Seq`; `toBWTW8` is `unpackBWT . toBWTPackedW8`.
data PackedBWT = PackedBWT { bwtColumn :: !BS.ByteString, bwtPrimary :: !Int }

toBWTPackedW8 :: BS.ByteString -> PackedBWT
toBWTPackedW8 xs
  | BS.null xs = PackedBWT BS.empty 0
  | otherwise  = unsafePerformIO $ withB200 $ \ctx ->
      BSU.unsafeUseAsCStringLen xs $ \(p, n) -> do
        bwt <- pinned (n + 1)
        alloca $ \pprim -> withForeignPtr bwt $ \pb -> do
          c_bwt_encode ctx (castPtr p) (fromIntegral n) pb pprim nullPtr >>= check ctx
          prim <- fromIntegral <$> peek pprim
          col  <- BS.packCStringLen (castPtr pb, n + 1)
          pure (PackedBWT col prim)

-- | Inverse of `toBWTPackedW8` (tc_bwt_decode_u8: the column as bytes + primary, 1 byte per symbol on the device).
fromBWTPackedW8 :: PackedBWT -> BS.ByteString
fromBWTPackedW8 (PackedBWT col prim)
  | BS.null col = BS.empty
  | otherwise   = unsafePerformIO $ withB200 $ \ctx ->
      BSU.unsafeUseAsCStringLen col $ \(p, bigN) -> do
        out <- pinned bigN
        alloca $ \pn -> withForeignPtr out $ \po -> do
          c_bwt_decode_u8 ctx (castPtr p) (fromIntegral bigN) (fromIntegral prim) po (fromIntegral bigN) pn >>= check ctx
          m <- fromIntegral <$> peek pn
          BS.packCStringLen (castPtr po, m)

-- | Move-to-front of a packed BWT (tc_mtf_encode_u8): the N indices as little-endian Word16 pairs of bytes in one
-- strict ByteString (2 N bytes; index k is bytes 2k, 2k + 1) and the final list.
toMTFPackedW8 :: PackedBWT -> (BS.ByteString, Seq (Maybe Word8))
toMTFPackedW8 (PackedBWT col prim)
  | BS.null col = (BS.empty, DS.empty)
  | otherwise   = unsafePerformIO $ withB200 $ \ctx ->
      BSU.unsafeUseAsCStringLen col $ \(p, bigN) -> do
        idx <- pinned (2 * bigN)
        fin <- pinned (2 * 257)
        alloca $ \psig -> withForeignPtr idx $ \px -> withForeignPtr fin $ \pf -> do
          c_mtf_encode_u8 ctx (castPtr p) (fromIntegral bigN) (fromIntegral prim) px pf psig >>= check ctx
          sg <- fromIntegral <$> peek psig
          fs <- peekArray sg pf
          ix <- BS.packCStringLen (castPtr px, 2 * bigN)
          pure (ix, DS.fromList (map i16ToMaybe fs))

-- | Device-resident FM-index handle (tc_fm); freed by the GC finaliser.
newtype B200FM = B200FM (ForeignPtr TcFm)

-- | *ToBWTToFMIndex* (Data/FMIndex.hs:108-183); sa_sample_rate 1 keeps the full SA.
buildFMIndexW8 :: Int -> BS.ByteString -> B200FM
buildFMIndexW8 rate xs = unsafePerformIO $ withB200 $ \ctx ->
  BSU.unsafeUseAsCStringLen xs $ \(p, n) -> alloca $ \pp -> do
    c_fm_build ctx (castPtr p) (fromIntegral n) (fromIntegral rate) pp >>= check ctx
    B200FM <$> (peek pp >>= newForeignPtr p_fm_free)

packPatterns :: [BS.ByteString] -> (BS.ByteString, [Word64])
packPatterns ps = (BS.concat ps, scanl (+) 0 (map (fromIntegral . BS.length) ps))

-- | countFMIndex (Data/FMIndex/Internal.hs:347-438) for a batch; input order is kept.
countFMIndexW8 :: B200FM -> [BS.ByteString] -> [Maybe Int]
countFMIndexW8 (B200FM fm) pats = unsafePerformIO $ withB200 $ \ctx -> withForeignPtr fm $ \pfm -> do
  let (flat, offs) = packPatterns pats
      q = length pats
  off <- pinned (8 * (q + 1))
  out <- pinned (8 * max 1 q)
  BSU.unsafeUseAsCStringLen flat $ \(pp, _) -> withForeignPtr off $ \po -> withForeignPtr out $ \pc -> do
    pokeArray po offs
    c_fm_count ctx pfm (castPtr pp) po (fromIntegral q) pc >>= check ctx
    map (\c -> if c < 0 then Nothing else Just (fromIntegral c)) <$> (peekArray q pc :: IO [Int64])

-- | locateFMIndex (Data/FMIndex/Internal.hs:448-542) + the rank -> suffixstartpos map of the
-- wrappers (Data/FMIndex.hs:496): 1-based positions in SA-rank order, per pattern.
locateFMIndexW8 :: B200FM -> [BS.ByteString] -> [Seq (Maybe Int)]
locateFMIndexW8 (B200FM fm) pats = unsafePerformIO $ withB200 $ \ctx -> withForeignPtr fm $ \pfm -> do
  let (flat, offs) = packPatterns pats
      q = length pats
  off  <- pinned (8 * (q + 1))
  hoff <- pinned (8 * (q + 1))
  BSU.unsafeUseAsCStringLen flat $ \(pp, _) -> withForeignPtr off $ \po -> withForeignPtr hoff $ \ph ->
    alloca $ \pt -> do
      pokeArray po offs
      rc <- c_fm_locate ctx pfm (castPtr pp) po (fromIntegral q) ph nullPtr 0 pt   -- sizing call
      when (rc /= 0 && rc /= tcECap) $ check ctx rc
      total <- fromIntegral <$> peek pt
      pos <- pinned (8 * max 1 total)
      withForeignPtr pos $ \ppos -> do
        c_fm_locate ctx pfm (castPtr pp) po (fromIntegral q) ph ppos (fromIntegral total) pt >>= check ctx
        hs <- map fromIntegral <$> (peekArray (q + 1) ph :: IO [Word64])
        ps <- map fromIntegral <$> (peekArray total ppos :: IO [Word64])
        pure (slices hs (DS.fromList (map Just ps)))

-- | Consecutive slices @[o0, o1), [o1, o2), ...@ of a sequence: O(log) per slice, not O(total).
slices :: [Int] -> Seq a -> [Seq a]
slices offs s = zipWith (\a b -> DS.take (b - a) (DS.drop a s)) offs (drop 1 offs)

-- | 'compressBlocksPackedW8' over all GPUs of the box: block b runs on device b `mod` (number of devices).
compressBlocksPackedMultiW8 :: [BS.ByteString] -> [BS.ByteString]
compressBlocksPackedMultiW8 []     = []
compressBlocksPackedMultiW8 blocks = unsafePerformIO $ do
  let nb   = length blocks
      ns   = map (fromIntegral . BS.length) blocks :: [Word64]
      devs = [0 .. fromIntegral (max 1 deviceCount) - 1] :: [CInt]
  caps <- mapM c_packed_bound ns
  outs <- mapM (pinned . fromIntegral) caps :: IO [ForeignPtr Word8]
  info <- pinned (552 * nb)
  withMany BSU.unsafeUseAsCStringLen blocks $ \ins ->
    withMany withForeignPtr outs $ \pouts ->
      withArray (map (castPtr . fst) ins) $ \ptext -> withArray devs $ \pdev ->
        withArray ns $ \pn -> withArray pouts $ \pout -> withArray caps $ \pcap ->
          withArray (replicate nb 0) $ \pbytes -> withForeignPtr info $ \pinfo -> do
            c_mgpu_blocks_encode_packed (fromIntegral (length devs)) pdev (fromIntegral nb) ptext pn 1 pout pcap pbytes pinfo
              >>= checkRc
            sizes <- peekArray nb pbytes
            forM (zip pouts sizes) $ \(po, sz) -> BS.packCStringLen (castPtr po, fromIntegral sz)

-- | One replica of an index per GPU (tc_fm_replicate: peer copies of the image over NVLink).
data B200FMs = B200FMs [CInt] [ForeignPtr TcFm]

replicateFMIndexW8 :: B200FM -> B200FMs
replicateFMIndexW8 (B200FM root) = unsafePerformIO $ withForeignPtr root $ \proot -> do
  let devs = [0 .. fromIntegral (max 1 deviceCount) - 1] :: [CInt]
      nd   = length devs
  withArray devs $ \pdev -> withArray (replicate nd nullPtr) $ \prep -> do
    c_fm_replicate proot (fromIntegral nd) pdev prep >>= checkRc
    reps <- peekArray nd prep
    B200FMs devs <$> mapM (newForeignPtr p_fm_free) reps

-- | 'countFMIndexW8' with the patterns split into contiguous chunks, one per GPU; input order is kept.
countFMIndexMultiW8 :: B200FMs -> [BS.ByteString] -> [Maybe Int]
countFMIndexMultiW8 (B200FMs devs reps) pats = unsafePerformIO $ do
  let (flat, offs) = packPatterns pats
      q = length pats
  off <- pinned (8 * (q + 1))
  out <- pinned (8 * max 1 q)
  withMany withForeignPtr reps $ \preps -> withArray preps $ \prep -> withArray devs $ \pdev ->
    BSU.unsafeUseAsCStringLen flat $ \(pp, _) -> withForeignPtr off $ \po -> withForeignPtr out $ \pc -> do
      pokeArray po offs
      c_mgpu_fm_count (fromIntegral (length devs)) pdev prep (castPtr pp) po (fromIntegral q) pc >>= checkRc
      map (\c -> if c < 0 then Nothing else Just (fromIntegral c)) <$> (peekArray q pc :: IO [Int64])

-- | 'locateFMIndexW8' over all GPUs.
locateFMIndexMultiW8 :: B200FMs -> [BS.ByteString] -> [Seq (Maybe Int)]
locateFMIndexMultiW8 (B200FMs devs reps) pats = unsafePerformIO $ do
  let (flat, offs) = packPatterns pats
      q  = length pats
      nd = fromIntegral (length devs)
  off  <- pinned (8 * (q + 1))
  hoff <- pinned (8 * (q + 1))
  withMany withForeignPtr reps $ \preps -> withArray preps $ \prep -> withArray devs $ \pdev ->
    BSU.unsafeUseAsCStringLen flat $ \(pp, _) -> withForeignPtr off $ \po -> withForeignPtr hoff $ \ph ->
      alloca $ \pt -> do
        pokeArray po offs
        rc <- c_mgpu_fm_locate nd pdev prep (castPtr pp) po (fromIntegral q) ph nullPtr 0 pt   -- sizing call
        when (rc /= 0 && rc /= tcECap) $ checkRc rc
        total <- fromIntegral <$> peek pt
        pos <- pinned (8 * max 1 total)
        withForeignPtr pos $ \ppos -> do
          c_mgpu_fm_locate nd pdev prep (castPtr pp) po (fromIntegral q) ph ppos (fromIntegral total) pt >>= checkRc
          hs <- map fromIntegral <$> (peekArray (q + 1) ph :: IO [Word64])
          ps <- map fromIntegral <$> (peekArray total ppos :: IO [Word64])
          pure (slices hs (DS.fromList (map Just ps)))
