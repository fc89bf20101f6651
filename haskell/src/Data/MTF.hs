{-# LANGUAGE OverloadedStrings #-}

-- |
-- Module      :  Data.MTF
-- Description :  drop-in replacement of text-compression's Data.MTF over the B200 kernels
--
-- Export list and types of the reference (src/Data/MTF.hs:36-63).
-- NOT COMPILED: no GHC exists in the build image (see "Data.TextCompression.B200").
module Data.MTF ( -- * To MTF functions
                  bytestringToBWTToMTFB,
                  bytestringToBWTToMTFT,
                  textToBWTToMTFB,
                  textToBWTToMTFT,
                  textBWTToMTFB,
                  bytestringBWTToMTFB,
                  textBWTToMTFT,
                  bytestringBWTToMTFT,
                  textToMTFB,
                  bytestringToMTFB,
                  textToMTFT,
                  bytestringToMTFT,
                  -- * From MTF functions
                  bytestringFromBWTFromMTFB,
                  bytestringFromBWTFromMTFT,
                  textFromBWTFromMTFB,
                  textFromBWTFromMTFT,
                  textBWTFromMTFT,
                  bytestringBWTFromMTFT,
                  textBWTFromMTFB,
                  bytestringBWTFromMTFB,
                  textFromMTFB,
                  bytestringFromMTFB,
                  textFromMTFT,
                  bytestringFromMTFT,
                  tests
                ) where

import           Data.BWT           hiding (tests)
import           Data.BWT.Internal
import           Data.MTF.Internal

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

mapList :: (a -> b) -> (Seq Int,Seq (Maybe a)) -> MTF b
mapList f (is, fl) = MTF (is, fmap (fmap f) fl)

{- to MTF -}

bytestringToBWTToMTFB :: ByteString -> MTF ByteString
bytestringToBWTToMTFB = bytestringBWTToMTFB . bytestringToBWT

bytestringToBWTToMTFT :: ByteString -> MTF Text
bytestringToBWTToMTFT = bytestringBWTToMTFT . bytestringToBWT

textToBWTToMTFB :: Text -> MTF ByteString
textToBWTToMTFB = textBWTToMTFB . textToBWT

textToBWTToMTFT :: Text -> MTF Text
textToBWTToMTFT = textBWTToMTFT . textToBWT

textBWTToMTFB :: TextBWT -> MTF ByteString
textBWTToMTFB = MTF . seqToMTF . fmap (fmap byteB) . unTextBWT

bytestringBWTToMTFB :: BWT Word8 -> MTF ByteString
bytestringBWTToMTFB (BWT xs) = MTF (seqToMTF (fmap (fmap byteB) xs))

textBWTToMTFT :: TextBWT -> MTF Text
textBWTToMTFT = MTF . seqToMTF . fmap (fmap byteT) . unTextBWT

bytestringBWTToMTFT :: BWT Word8 -> MTF Text
bytestringBWTToMTFT (BWT xs) = MTF (seqToMTF (fmap (fmap byteT) xs))

textToMTFB :: Seq (Maybe Text) -> MTF ByteString
textToMTFB DS.Empty = MTF (DS.Empty,DS.Empty)
textToMTFB xs       = MTF (seqToMTF (fmap (fmap DTE.encodeUtf8) xs))

bytestringToMTFB :: Seq (Maybe ByteString) -> MTF ByteString
bytestringToMTFB DS.Empty = MTF (DS.Empty,DS.Empty)
bytestringToMTFB xs       = MTF (seqToMTF xs)

textToMTFT :: Seq (Maybe Text) -> MTF Text
textToMTFT DS.Empty = MTF (DS.Empty,DS.Empty)
textToMTFT xs       = MTF (seqToMTF xs)

bytestringToMTFT :: Seq (Maybe ByteString) -> MTF Text
bytestringToMTFT DS.Empty = MTF (DS.Empty,DS.Empty)
bytestringToMTFT xs       = MTF (seqToMTF (fmap (fmap DTE.decodeUtf8) xs))

{- from MTF -}

bytestringFromBWTFromMTFB :: MTF ByteString -> ByteString
bytestringFromBWTFromMTFB = bytestringFromByteStringBWT . bytestringBWTFromMTFB

bytestringFromBWTFromMTFT :: MTF Text -> ByteString
bytestringFromBWTFromMTFT = bytestringFromByteStringBWT . bytestringBWTFromMTFT

textFromBWTFromMTFB :: MTF ByteString -> Text
textFromBWTFromMTFB = DTE.decodeUtf8 . bytestringFromByteStringBWT . bytestringBWTFromMTFB

textFromBWTFromMTFT :: MTF Text -> Text
textFromBWTFromMTFT = DTE.decodeUtf8 . bytestringFromByteStringBWT . bytestringBWTFromMTFT

textBWTFromMTFT :: MTF Text -> BWT Text
textBWTFromMTFT = BWT . seqFromMTF

bytestringBWTFromMTFT :: MTF Text -> BWT ByteString
bytestringBWTFromMTFT = BWT . fmap (fmap DTE.encodeUtf8) . seqFromMTF

textBWTFromMTFB :: MTF ByteString -> BWT Text
textBWTFromMTFB = BWT . fmap (fmap DTE.decodeUtf8) . seqFromMTF

bytestringBWTFromMTFB :: MTF ByteString -> BWT ByteString
bytestringBWTFromMTFB = BWT . seqFromMTF

textFromMTFB :: MTF ByteString -> Seq (Maybe Text)
textFromMTFB = fmap (fmap DTE.decodeUtf8) . seqFromMTF

bytestringFromMTFB :: MTF ByteString -> Seq (Maybe ByteString)
bytestringFromMTFB = seqFromMTF

textFromMTFT :: MTF Text -> Seq (Maybe Text)
textFromMTFT = seqFromMTF

bytestringFromMTFT :: MTF Text -> Seq (Maybe ByteString)
bytestringFromMTFT = fmap (fmap DTE.encodeUtf8) . seqFromMTF

{- tests -}

-- | The two known-answer vectors of the reference (src/Data/MTF.hs:287-299).
tests :: Test
tests = TestList
  [ TestLabel "to MTF"   (TestCase (assertEqual "aaabbbccc" mtf1 (textToBWTToMTFB t1)))
  , TestLabel "from MTF" (TestCase (assertEqual "aaabbbccc" t1 (textFromBWTFromMTFB mtf1)))
  ]
  where
    t1   = DTE.decodeUtf8 "aaabbbccc"
    mtf1 = mapList id (DS.fromList [3,1,2,0,0,3,0,3,0,1], DS.fromList [Just "b", Just "c", Just "a", Nothing])
