{-# LANGUAGE OverloadedStrings #-}
-- |
-- Module      :  Data.RLE
-- Description :  drop-in replacement of text-compression's Data.RLE over the B200 kernels
--
-- Export list and types of the reference (src/Data/RLE.hs:35-62).  The @...ToBWTToRLE...@ helpers run
-- BWT and run-length encoding on the device.  The T (Text element) variants decode every byte on its own, as
-- the reference does, and therefore throw on bytes >= 0x80 (SURVEY.md 2.3 Q6).
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.RLE ( -- * To RLE functions
                  bytestringToBWTToRLEB,
                  bytestringToBWTToRLET,
                  textToBWTToRLEB,
                  textToBWTToRLET,
                  textBWTToRLEB,
                  bytestringBWTToRLEB,
                  textBWTToRLET,
                  bytestringBWTToRLET,
                  textToRLEB,
                  bytestringToRLEB,
                  textToRLET,
                  bytestringToRLET,
                  -- * From RLE functions
                  bytestringFromBWTFromRLEB,
                  bytestringFromBWTFromRLET,
                  textFromBWTFromRLEB,
                  textFromBWTFromRLET,
                  textBWTFromRLET,
                  bytestringBWTFromRLET,
                  textBWTFromRLEB,
                  bytestringBWTFromRLEB,
                  textFromRLEB,
                  bytestringFromRLEB,
                  textFromRLET,
                  bytestringFromRLET,
                  tests
                ) where

import           Data.BWT           hiding (tests)
import           Data.BWT.Internal
import           Data.RLE.Internal

import           Data.ByteString    (ByteString)
import qualified Data.ByteString    as BS
import           Data.Sequence      (Seq (..))
import qualified Data.Sequence      as DS
import           Data.Text          (Text)
import qualified Data.Text.Encoding as DTE
import           Data.Word          (Word8)
import           Test.HUnit

byteB :: Word8 -> ByteString
byteB = BS.singleton

byteT :: Word8 -> Text
byteT = DTE.decodeUtf8 . BS.singleton

unTextBWT :: TextBWT -> Seq (Maybe Word8)
unTextBWT (TextBWT (BWT xs)) = xs

{- to RLE -}

bytestringToBWTToRLEB :: ByteString -> RLE ByteString
bytestringToBWTToRLEB = bytestringBWTToRLEB . bytestringToBWT

bytestringToBWTToRLET :: ByteString -> RLE Text
bytestringToBWTToRLET = bytestringBWTToRLET . bytestringToBWT

textToBWTToRLEB :: Text -> RLE ByteString
textToBWTToRLEB = textBWTToRLEB . textToBWT

textToBWTToRLET :: Text -> RLE Text
textToBWTToRLET = textBWTToRLET . textToBWT

textBWTToRLEB :: TextBWT -> RLE ByteString
textBWTToRLEB = RLE . seqToRLE . fmap (fmap byteB) . unTextBWT

bytestringBWTToRLEB :: BWT Word8 -> RLE ByteString
bytestringBWTToRLEB (BWT xs) = RLE (seqToRLE (fmap (fmap byteB) xs))

textBWTToRLET :: TextBWT -> RLE Text
textBWTToRLET = RLE . seqToRLE . fmap (fmap byteT) . unTextBWT

bytestringBWTToRLET :: BWT Word8 -> RLE Text
bytestringBWTToRLET (BWT xs) = RLE (seqToRLE (fmap (fmap byteT) xs))

textToRLEB :: Seq (Maybe Text) -> RLE ByteString
textToRLEB = RLE . seqToRLE . fmap (fmap DTE.encodeUtf8)

bytestringToRLEB :: Seq (Maybe ByteString) -> RLE ByteString
bytestringToRLEB = RLE . seqToRLE

textToRLET :: Seq (Maybe Text) -> RLE Text
textToRLET = RLE . seqToRLE

bytestringToRLET :: Seq (Maybe ByteString) -> RLE Text
bytestringToRLET = RLE . seqToRLE . fmap (fmap DTE.decodeUtf8)

{- from RLE -}

bytestringFromBWTFromRLEB :: RLE ByteString -> ByteString
bytestringFromBWTFromRLEB = bytestringFromByteStringBWT . bytestringBWTFromRLEB

bytestringFromBWTFromRLET :: RLE Text -> ByteString
bytestringFromBWTFromRLET = bytestringFromByteStringBWT . bytestringBWTFromRLET

textFromBWTFromRLEB :: RLE ByteString -> Text
textFromBWTFromRLEB = DTE.decodeUtf8 . bytestringFromByteStringBWT . bytestringBWTFromRLEB

textFromBWTFromRLET :: RLE Text -> Text
textFromBWTFromRLET = DTE.decodeUtf8 . bytestringFromByteStringBWT . bytestringBWTFromRLET

textBWTFromRLET :: RLE Text -> BWT Text
textBWTFromRLET = BWT . seqFromRLE

bytestringBWTFromRLET :: RLE Text -> BWT ByteString
bytestringBWTFromRLET = BWT . fmap (fmap DTE.encodeUtf8) . seqFromRLE

textBWTFromRLEB :: RLE ByteString -> BWT Text
textBWTFromRLEB = BWT . fmap (fmap DTE.decodeUtf8) . seqFromRLE

bytestringBWTFromRLEB :: RLE ByteString -> BWT ByteString
bytestringBWTFromRLEB = BWT . seqFromRLE

textFromRLEB :: RLE ByteString -> Seq (Maybe Text)
textFromRLEB = fmap (fmap DTE.decodeUtf8) . seqFromRLE

bytestringFromRLEB :: RLE ByteString -> Seq (Maybe ByteString)
bytestringFromRLEB = seqFromRLE

textFromRLET :: RLE Text -> Seq (Maybe Text)
textFromRLET = seqFromRLE

bytestringFromRLET :: RLE Text -> Seq (Maybe ByteString)
bytestringFromRLET = fmap (fmap DTE.encodeUtf8) . seqFromRLE

{- tests -}

-- | The first known-answer vector of the reference (src/Data/RLE.hs:279-288,316,318) and the round trip;
-- all six reference vectors are replayed through the C ABI in tests/test_gpu_parity.py.
tests :: Test
tests = TestList
  [ TestLabel "to RLE"   (TestCase (assertEqual "aaaabbbbcccc" rle1 (textToBWTToRLET s1)))
  , TestLabel "from RLE" (TestCase (assertEqual "aaaabbbbcccc" s1 (textFromBWTFromRLET rle1)))
  ]
  where
    s1   = DTE.decodeUtf8 "aaaabbbbcccc"
    j    = Just . DTE.decodeUtf8
    rle1 = RLE (DS.fromList [j "1", j "c", j "1", Nothing, j "4", j "a", j "3", j "b", j "3", j "c", j "1", j "b"])
