-- |
-- Module      :  Data.FMIndex
-- Description :  drop-in replacement of text-compression's Data.FMIndex over the B200 kernels
--
-- Export list and types of the reference (src/Data/FMIndex.hs:49-82).  The batch count / locate functions build
-- the device-resident index once per call (the reference rebuilds its index per call as well) and answer all
-- patterns with one kernel launch; the ...P variants split the patterns over every GPU of the box the way the
-- reference splits them over the cores with parListChunk (src/Data/FMIndex.hs:417-423, 544-553).  Positions are
-- 1-based and come in suffix-array order, not sorted, like the reference's.
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.FMIndex ( -- * To FMIndex functions
                      bytestringToBWTToFMIndexB,
                      bytestringToBWTToFMIndexT,
                      textToBWTToFMIndexB,
                      textToBWTToFMIndexT,
                      textBWTToFMIndexB,
                      bytestringBWTToFMIndexB,
                      textBWTToFMIndexT,
                      bytestringBWTToFMIndexT,
                      -- * From FMIndex functions
                      bytestringFromBWTFromFMIndexB,
                      bytestringFromBWTFromFMIndexT,
                      textFromBWTFromFMIndexB,
                      textFromBWTFromFMIndexT,
                      textBWTFromFMIndexT,
                      bytestringBWTFromFMIndexT,
                      textBWTFromFMIndexB,
                      bytestringBWTFromFMIndexB,
                      textFromFMIndexB,
                      bytestringFromFMIndexB,
                      textFromFMIndexT,
                      bytestringFromFMIndexT,
                      -- * Count operations
                      bytestringFMIndexCountS,
                      textFMIndexCountS,
                      bytestringFMIndexCountP,
                      textFMIndexCountP,
                      -- * Locate operations
                      bytestringFMIndexLocateS,
                      textFMIndexLocateS,
                      bytestringFMIndexLocateP,
                      textFMIndexLocateP,
                      tests
                    ) where

import           Data.BWT                  hiding (tests)
import           Data.BWT.Internal
import           Data.FMIndex.Internal     (Cc (Cc), FMIndex (FMIndex), OccCK (OccCK), SA (SA), seqFromFMIndex,
                                            seqToCc, seqToOccCK)

import           Data.ByteString           (ByteString)
import qualified Data.ByteString           as BS
import           Data.Sequence             (Seq (..))
import qualified Data.Sequence             as DS
import           Data.Text                 (Text)
import qualified Data.Text.Encoding        as DTE
import qualified Data.TextCompression.B200 as B200
import           Data.Word                 (Word8)
import           Test.HUnit

byteB :: Word8 -> ByteString
byteB = BS.singleton

byteT :: Word8 -> Text
byteT = DTE.decodeUtf8 . BS.singleton

-- | SA rate of the device index the batch functions build: every 32nd text position keeps its suffix-array
-- entry, the others are reached by walking LF steps.
saRate :: Int
saRate = 32

{- to FMIndex -}

-- the first column of the sorted rotations = the sorted symbols of text$
firstColumn :: BWTMatrix Word8 -> Seq (Maybe Word8)
firstColumn (BWTMatrix rows) = fmap (\r -> DS.index r 0) rows

assemble :: (Ord b) => (Word8 -> b) -> (Seq (Maybe b) -> Seq (Maybe b,Seq (Int,Int,Maybe b)))
         -> BWTMatrix Word8 -> BWT Word8 -> SuffixArray b -> FMIndex b
assemble f occOf m (BWT col) sa =
  FMIndex (Cc (seqToCc (fmap (fmap f) (firstColumn m))), OccCK (occOf (fmap (fmap f) col)), SA sa)

-- the suffix array the index stores is that of the text the BWT decodes to (src/Data/FMIndex.hs:143-147)
suffixesOf :: Ord b => (Word8 -> b) -> BWT Word8 -> SuffixArray b
suffixesOf f bwt = createSuffixArray (DS.fromList (map f (BS.unpack (bytestringFromWord8BWT bwt))))

bytestringToBWTToFMIndexB :: ByteString -> FMIndex ByteString
bytestringToBWTToFMIndexB xs = bytestringBWTToFMIndexB (createBWTMatrix (BS.unpack xs)) (bytestringToBWT xs)

bytestringToBWTToFMIndexT :: ByteString -> FMIndex Text
bytestringToBWTToFMIndexT xs = bytestringBWTToFMIndexT (createBWTMatrix (BS.unpack xs)) (bytestringToBWT xs)

textToBWTToFMIndexB :: Text -> FMIndex ByteString
textToBWTToFMIndexB xs = textBWTToFMIndexB (createBWTMatrix (BS.unpack (DTE.encodeUtf8 xs))) (textToBWT xs)

textToBWTToFMIndexT :: Text -> FMIndex Text
textToBWTToFMIndexT xs = textBWTToFMIndexT (createBWTMatrix (BS.unpack (DTE.encodeUtf8 xs))) (textToBWT xs)

textBWTToFMIndexB :: BWTMatrix Word8 -> TextBWT -> FMIndex ByteString
textBWTToFMIndexB (BWTMatrix DS.Empty) _ = FMIndex (Cc DS.Empty,OccCK DS.Empty,SA DS.Empty)
textBWTToFMIndexB m (TextBWT bwt)        = assemble byteB seqToOccCK m bwt (suffixesOf byteB bwt)

bytestringBWTToFMIndexB :: BWTMatrix Word8 -> BWT Word8 -> FMIndex ByteString
bytestringBWTToFMIndexB (BWTMatrix DS.Empty) _ = FMIndex (Cc DS.Empty,OccCK DS.Empty,SA DS.Empty)
bytestringBWTToFMIndexB m bwt                  = assemble byteB seqToOccCK m bwt (suffixesOf byteB bwt)

textBWTToFMIndexT :: BWTMatrix Word8 -> TextBWT -> FMIndex Text
textBWTToFMIndexT (BWTMatrix DS.Empty) _ = FMIndex (Cc DS.Empty,OccCK DS.Empty,SA DS.Empty)
textBWTToFMIndexT m (TextBWT bwt)        = assemble byteT seqToOccCK m bwt (suffixesOf byteT bwt)

bytestringBWTToFMIndexT :: BWTMatrix Word8 -> BWT Word8 -> FMIndex Text
bytestringBWTToFMIndexT (BWTMatrix DS.Empty) _ = FMIndex (Cc DS.Empty,OccCK DS.Empty,SA DS.Empty)
bytestringBWTToFMIndexT m bwt                  = assemble byteT seqToOccCK m bwt (suffixesOf byteT bwt)

{- from FMIndex -}

bytestringFromBWTFromFMIndexB :: FMIndex ByteString -> ByteString
bytestringFromBWTFromFMIndexB = bytestringFromByteStringBWT . bytestringBWTFromFMIndexB

bytestringFromBWTFromFMIndexT :: FMIndex Text -> ByteString
bytestringFromBWTFromFMIndexT = bytestringFromByteStringBWT . bytestringBWTFromFMIndexT

textFromBWTFromFMIndexB :: FMIndex ByteString -> Text
textFromBWTFromFMIndexB = DTE.decodeUtf8 . bytestringFromByteStringBWT . bytestringBWTFromFMIndexB

textFromBWTFromFMIndexT :: FMIndex Text -> Text
textFromBWTFromFMIndexT = DTE.decodeUtf8 . bytestringFromByteStringBWT . bytestringBWTFromFMIndexT

textBWTFromFMIndexT :: FMIndex Text -> BWT Text
textBWTFromFMIndexT = BWT . seqFromFMIndex

bytestringBWTFromFMIndexT :: FMIndex Text -> BWT ByteString
bytestringBWTFromFMIndexT = BWT . fmap (fmap DTE.encodeUtf8) . seqFromFMIndex

textBWTFromFMIndexB :: FMIndex ByteString -> BWT Text
textBWTFromFMIndexB = BWT . fmap (fmap DTE.decodeUtf8) . seqFromFMIndex

bytestringBWTFromFMIndexB :: FMIndex ByteString -> BWT ByteString
bytestringBWTFromFMIndexB = BWT . seqFromFMIndex

textFromFMIndexB :: FMIndex ByteString -> Seq (Maybe Text)
textFromFMIndexB = fmap (fmap DTE.decodeUtf8) . seqFromFMIndex

bytestringFromFMIndexB :: FMIndex ByteString -> Seq (Maybe ByteString)
bytestringFromFMIndexB = seqFromFMIndex

textFromFMIndexT :: FMIndex Text -> Seq (Maybe Text)
textFromFMIndexT = seqFromFMIndex

bytestringFromFMIndexT :: FMIndex Text -> Seq (Maybe ByteString)
bytestringFromFMIndexT = fmap (fmap DTE.encodeUtf8) . seqFromFMIndex

{- count -}

bytestringFMIndexCountS :: [ByteString] -> ByteString -> Seq (ByteString,Maybe Int)
bytestringFMIndexCountS []   _     = DS.Empty
bytestringFMIndexCountS pats input
  | BS.null input = DS.Empty
  | otherwise     = DS.fromList (zip pats (B200.countFMIndexW8 (B200.buildFMIndexW8 saRate input) pats))

textFMIndexCountS :: [Text] -> Text -> Seq (Text,Maybe Int)
textFMIndexCountS pats input =
  DS.zipWith (\p (_, c) -> (p, c)) (DS.fromList pats)
             (bytestringFMIndexCountS (map DTE.encodeUtf8 pats) (DTE.encodeUtf8 input))

-- | As the S variant, with the patterns split into one contiguous chunk per GPU.
bytestringFMIndexCountP :: [ByteString] -> ByteString -> IO (Seq (ByteString,Maybe Int))
bytestringFMIndexCountP []   _     = return DS.Empty
bytestringFMIndexCountP pats input
  | BS.null input = return DS.Empty
  | otherwise     = do
      let replicas = B200.replicateFMIndexW8 (B200.buildFMIndexW8 saRate input)
      return $! DS.fromList (zip pats (B200.countFMIndexMultiW8 replicas pats))

textFMIndexCountP :: [Text] -> Text -> IO (Seq (Text,Maybe Int))
textFMIndexCountP pats input = do
  r <- bytestringFMIndexCountP (map DTE.encodeUtf8 pats) (DTE.encodeUtf8 input)
  return (DS.zipWith (\p (_, c) -> (p, c)) (DS.fromList pats) r)

{- locate -}

bytestringFMIndexLocateS :: [ByteString] -> ByteString -> Seq (ByteString,Seq (Maybe Int))
bytestringFMIndexLocateS []   _     = DS.Empty
bytestringFMIndexLocateS pats input
  | BS.null input = DS.Empty
  | otherwise     = DS.fromList (zip pats (B200.locateFMIndexW8 (B200.buildFMIndexW8 saRate input) pats))

textFMIndexLocateS :: [Text] -> Text -> Seq (Text,Seq (Maybe Int))
textFMIndexLocateS pats input =
  DS.zipWith (\p (_, l) -> (p, l)) (DS.fromList pats)
             (bytestringFMIndexLocateS (map DTE.encodeUtf8 pats) (DTE.encodeUtf8 input))

bytestringFMIndexLocateP :: [ByteString] -> ByteString -> IO (Seq (ByteString,Seq (Maybe Int)))
bytestringFMIndexLocateP []   _     = return DS.Empty
bytestringFMIndexLocateP pats input
  | BS.null input = return DS.Empty
  | otherwise     = do
      let replicas = B200.replicateFMIndexW8 (B200.buildFMIndexW8 saRate input)
      return $! DS.fromList (zip pats (B200.locateFMIndexMultiW8 replicas pats))

textFMIndexLocateP :: [Text] -> Text -> IO (Seq (Text,Seq (Maybe Int)))
textFMIndexLocateP pats input = do
  r <- bytestringFMIndexLocateP (map DTE.encodeUtf8 pats) (DTE.encodeUtf8 input)
  return (DS.zipWith (\p (_, l) -> (p, l)) (DS.fromList pats) r)

-- | The reference has no FM-index test (src/Data/FMIndex.hs:603-604).
tests :: Test
tests = TestList []
